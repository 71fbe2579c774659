// TEST INFRASTRUCTURE ONLY (oracle build). Minimal stand-in for google-glog so that the
// reference headers under /root/reference/include compile offline.  The reference's CMake
// downloads glog (cmakemodules/DownloadGlog.cmake:1-6); there is no network here.
// Only the macros the reference uses are provided: CHECK*/DCHECK* (abort on failure /
// no-op under NDEBUG), LOG/DLOG/VLOG/DVLOG (discarded), and the FLAGS_* the tdc driver sets
// (include/tudocomp_driver/Options.hpp:239-249, src/tudocomp_driver/tudocomp_driver.cpp:58,86).
#pragma once
#include <cassert>
#include <climits>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

namespace tdc_glog_shim {
struct NullStream {
    template <class T> NullStream& operator<<(const T&) { return *this; }
    NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct FatalStream {
    std::ostringstream ss;
    FatalStream(const char* file, int line, const char* what) { ss << file << ":" << line << ": CHECK failed: " << what << " "; }
    template <class T> FatalStream& operator<<(const T& v) { ss << v; return *this; }
    FatalStream& operator<<(std::ostream& (*f)(std::ostream&)) { ss << f; return *this; }
    [[noreturn]] ~FatalStream() { std::cerr << ss.str() << std::endl; std::abort(); }
};
struct Voidify { template <class T> void operator&(T&&) {} };
}  // namespace tdc_glog_shim

#define TDC_SHIM_NULL() ::tdc_glog_shim::NullStream()
#define TDC_SHIM_CHECK(cond, text) \
    (cond) ? (void)0 : ::tdc_glog_shim::Voidify() & ::tdc_glog_shim::FatalStream(__FILE__, __LINE__, text)
#define TDC_SHIM_DEAD(cond) \
    true ? (void)0 : ::tdc_glog_shim::Voidify() & TDC_SHIM_NULL()

#define CHECK(c) TDC_SHIM_CHECK((c), #c)
#define CHECK_EQ(a, b) TDC_SHIM_CHECK((a) == (b), #a " == " #b)
#define CHECK_NE(a, b) TDC_SHIM_CHECK((a) != (b), #a " != " #b)
#define CHECK_LT(a, b) TDC_SHIM_CHECK((a) < (b), #a " < " #b)
#define CHECK_LE(a, b) TDC_SHIM_CHECK((a) <= (b), #a " <= " #b)
#define CHECK_GT(a, b) TDC_SHIM_CHECK((a) > (b), #a " > " #b)
#define CHECK_GE(a, b) TDC_SHIM_CHECK((a) >= (b), #a " >= " #b)

#ifdef NDEBUG
#define DCHECK(c) TDC_SHIM_DEAD(c)
#define DCHECK_EQ(a, b) TDC_SHIM_DEAD((a) == (b))
#define DCHECK_NE(a, b) TDC_SHIM_DEAD((a) != (b))
#define DCHECK_LT(a, b) TDC_SHIM_DEAD((a) < (b))
#define DCHECK_LE(a, b) TDC_SHIM_DEAD((a) <= (b))
#define DCHECK_GT(a, b) TDC_SHIM_DEAD((a) > (b))
#define DCHECK_GE(a, b) TDC_SHIM_DEAD((a) >= (b))
#else
#define DCHECK(c) CHECK(c)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) CHECK_NE(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#endif

// LOG evaluates its operands (glog would format and emit them); DLOG / DVLOG compile to dead code under NDEBUG exactly
// like glog's own definitions (`true ? (void)0 : LogMessageVoidify() & LOG(...)`), and VLOG is guarded by VLOG_IS_ON:
// their operands are NOT evaluated.  (An earlier version of this shim evaluated them, which made e.g.
// ChainCompressor's `DLOG(INFO) << vec_to_debug_string(between_buf)` format every intermediate buffer: 75 ns per byte and
// chain link that a real Release build of the reference does not pay.)
#define LOG(sev) TDC_SHIM_NULL()
#define LOG_IF(sev, c) TDC_SHIM_NULL()
#define VLOG_IS_ON(lvl) false
#define VLOG(lvl) TDC_SHIM_DEAD(0)
#ifdef NDEBUG
#define DLOG(sev) TDC_SHIM_DEAD(0)
#define DVLOG(lvl) TDC_SHIM_DEAD(0)
#else
#define DLOG(sev) TDC_SHIM_NULL()
#define DVLOG(lvl) TDC_SHIM_NULL()
#endif

static int FLAGS_logtostderr __attribute__((unused)) = 1;
static int FLAGS_v __attribute__((unused)) = 0;
static int FLAGS_minloglevel __attribute__((unused)) = 0;
static std::string FLAGS_log_dir __attribute__((unused));

namespace google {
inline void InitGoogleLogging(const char*) {}
}
