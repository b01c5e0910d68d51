// TEST INFRASTRUCTURE ONLY -- force-included compatibility header that lets the reference's vendored
// (MSVC/Windows/GSL-only) polar library, /root/reference/{headers,out}/external, compile with g++ so that
// its SC-list decoder can serve as the oracle for the polar rows.  Nothing here restates
// reference code: it only maps MSVC-isms onto their POSIX / ISO C++ equivalents.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <iostream>
#include <fstream>
#include <string>
#include <strings.h>
#include <type_traits>
// every standard header the library (or this shim) uses must be seen BEFORE the two macros at the end of this
// file (`exception`, `enable_if`) so that libstdc++ itself is compiled untouched
#include <chrono>
#include <complex>
#include <csetjmp>
#include <cstdint>
#include <ctime>
#include <functional>
#include <iomanip>
#include <istream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <ostream>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <utility>
#include <vector>
#include <immintrin.h>
#include <nmmintrin.h>

#define __int64 long long
#define __int32 int
#define __int16 short
#define __int8 char
#define _ASSERT(x) assert(x)
#define _CrtCheckMemory() 1
#define __declspec(x)
#define _stricmp strcasecmp
#define alloca __builtin_alloca
#ifndef _MAX_PATH
#define _MAX_PATH 4096
#endif

constexpr unsigned long long operator""ui64(unsigned long long v) { return v; }

inline void *_aligned_malloc(size_t size, size_t align) {
    void *p = nullptr;
    if (align < sizeof(void *)) align = sizeof(void *);
    if (posix_memalign(&p, align, size ? size : align)) return nullptr;
    return p;
}
inline void _aligned_free(void *p) { free(p); }
inline int vsprintf_s(char *buf, size_t n, const char *fmt, va_list ap) { return vsnprintf(buf, n, fmt, ap); }
inline int strcpy_s(char *dst, size_t n, const char *src) {
    strncpy(dst, src, n);
    if (n) dst[n - 1] = 0;
    return 0;
}
template <size_t N>
inline int strcpy_s(char (&dst)[N], const char *src) { return strcpy_s(dst, N, src); }
inline int sprintf_s(char *buf, size_t n, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    int r = vsnprintf(buf, n, fmt, ap);
    va_end(ap);
    return r;
}

// MSVC's std::exception has a (const char*) constructor and is assignable from it
namespace pk_compat {
class msvc_exception : public std::exception {
    std::string m_;
public:
    msvc_exception() {}
    msvc_exception(const char *m) : m_(m ? m : "") {}
    const char *what() const noexcept override { return m_.c_str(); }
};
}  // namespace pk_compat
namespace std {
using pk_msvc_exception = pk_compat::msvc_exception;
}
#define exception pk_msvc_exception

// MSVC accepts `typename std::enable_if<false>::type` as a default argument of a member template when the
// condition does not depend on that template (Simulation/Channel.h:91-108); g++ rejects it.  Those members are
// never instantiated by the decoder path, so an always-true stand-in is harmless.
namespace std {
template <bool B, class T = void>
struct pk_lenient_enable_if { typedef T type; };
}
#define enable_if pk_lenient_enable_if

using std::cerr;
using std::cout;
using std::endl;
using std::max;
using std::min;
using std::swap;
