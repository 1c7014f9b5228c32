// flopcount.cpp -- the oracle compiled with a COUNTING scalar type (the canonical
// algorithmic FLOP figure of one solve).  TEST/MEASUREMENT INFRASTRUCTURE.
// add / sub / mul / div / sqrt count 1 each (so a fused multiply-add counts 2); transcendental calls
// (sin cos exp log erf atan2 fmod) are counted separately.  Build: see Makefile (flops target).
#include <math.h>
#include <stdint.h>

struct Counted;
static thread_local uint64_t g_flops = 0, g_transc = 0;
struct Counted {
    double v;
    Counted() : v(0.0) {}
    Counted(double x) : v(x) {}
    Counted(int x) : v(x) {}
    explicit operator double() const { return v; }
    Counted& operator+=(const Counted& o) { g_flops++; v += o.v; return *this; }
    Counted& operator-=(const Counted& o) { g_flops++; v -= o.v; return *this; }
    Counted& operator*=(const Counted& o) { g_flops++; v *= o.v; return *this; }
};
static inline Counted operator+(const Counted& a, const Counted& b) { g_flops++; return Counted(a.v + b.v); }
static inline Counted operator-(const Counted& a, const Counted& b) { g_flops++; return Counted(a.v - b.v); }
static inline Counted operator*(const Counted& a, const Counted& b) { g_flops++; return Counted(a.v * b.v); }
static inline Counted operator/(const Counted& a, const Counted& b) { g_flops++; return Counted(a.v / b.v); }
static inline Counted operator-(const Counted& a) { return Counted(-a.v); }
static inline bool operator<(const Counted& a, const Counted& b) { return a.v < b.v; }
static inline bool operator>(const Counted& a, const Counted& b) { return a.v > b.v; }
static inline bool operator<=(const Counted& a, const Counted& b) { return a.v <= b.v; }
static inline bool operator>=(const Counted& a, const Counted& b) { return a.v >= b.v; }
static inline bool operator==(const Counted& a, const Counted& b) { return a.v == b.v; }
static inline bool operator!=(const Counted& a, const Counted& b) { return a.v != b.v; }
static inline Counted sqrt(const Counted& a) { g_flops++; return Counted(::sqrt(a.v)); }
static inline Counted fabs(const Counted& a) { return Counted(::fabs(a.v)); }
#define TR1(f) static inline Counted f(const Counted& a) { g_transc++; return Counted(::f(a.v)); }
TR1(sin) TR1(cos) TR1(exp) TR1(log) TR1(erf) TR1(tan) TR1(atan)
static inline Counted atan2(const Counted& a, const Counted& b) { g_transc++; return Counted(::atan2(a.v, b.v)); }
static inline Counted fmod(const Counted& a, const Counted& b) { g_transc++; return Counted(::fmod(a.v, b.v)); }
static_assert(sizeof(Counted) == sizeof(double), "Counted must alias double arrays");

#define REAL Counted
#define ORACLE_NO_API
#include "mpc_oracle.c"

extern "C" int flopcount_solve(int n, const double* xinit, const double* x0, const double* params, const int* num_iter,
                               double* flops, double* transc, int* ipm_iters, int* exit_code)
{
    setup_constraints();
    work_t* w = new work_t();
    for (int i = 0; i < n; i++) {
        static thread_local Counted xtraj[NX * (NN + 1)], utraj[NU * NN];
        Counted pobj, req;
        int ec, qs, ipm;
        g_flops = 0; g_transc = 0;
        solve_one(w, (const Counted*)xinit + (size_t)i * NX, (const Counted*)x0 + (size_t)i * NZ * (NN + 1),
                  (const Counted*)params + (size_t)i * NN * NP, num_iter[i], nullptr, xtraj, utraj, &pobj, &ec, &qs, &req, &ipm);
        flops[i] = (double)g_flops; transc[i] = (double)g_transc; ipm_iters[i] = ipm; exit_code[i] = ec;
    }
    delete w;
    return 0;
}
