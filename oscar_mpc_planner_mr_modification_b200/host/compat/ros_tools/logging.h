// Minimal stand-in for ros_tools/logging.h (external package, not vendored by the reference).
#pragma once
#include <iostream>
#define LOG_INFO(x) std::cout << x << std::endl
#define LOG_WARN(x) std::cerr << "[warn] " << x << std::endl
#define LOG_ERROR(x) std::cerr << "[error] " << x << std::endl
#define LOG_DEBUG(x) do { } while (0)
#define LOG_HEADER(x) std::cout << "== " << x << " ==" << std::endl
#define LOG_VALUE(n, v) std::cout << n << ": " << v << std::endl
#define LOG_WARN_THROTTLE(t, x) do { } while (0)
#define LOG_HOOK_MSG(x) do { } while (0)
#define LOG_INITIALIZED() do { } while (0)
