// Variants under test for tools/microbench/fac.cu.  Start as copies of the production routine; edit v1 / v2 to try an idea
// (this round: branch-free rsqrt 872 -> 745 cycles per stage, no operand gate in the Cholesky step -> 690, none in the fused
// G step -> 566, prefetch of the recursion-independent operands one stage ahead -> 660: rejected).
namespace MPC_NS {
__device__ __noinline__ void factor_v1(double* __restrict__ rs) { riccati_factor_coop(rs); }
__device__ __noinline__ void factor_v2(double* __restrict__ rs) { riccati_factor_coop(rs); }
}
