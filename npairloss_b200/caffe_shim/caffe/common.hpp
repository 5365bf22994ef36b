// Mini-Caffe shim: only what NPairMultiClassLossLayer touches.  NOT a Caffe re-implementation -- the reference is a
// leaf plugin inside a private Caffe fork that is not available (last row); this header exists so
// the layer sources compile and can be driven by tests.  Names and semantics follow BVLC Caffe's caffe/common.hpp plus
// the three fork statics the reference uses (Caffe::MULTI_GPU / NUM_GPU / RANK; npair_multi_class_loss.cpp:44-57,
// npair_multi_class_loss.cu:214,220,494).
#ifndef CAFFE_COMMON_HPP_
#define CAFFE_COMMON_HPP_

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace caffe {

using std::string;
using std::vector;
using std::shared_ptr;

// glog's LOG(FATAL)/CHECK abort the process.  The shim throws instead so that a test harness can observe the failure;
// nothing propagates across the C ABI (npairloss_b200/caffe_shim/harness.cpp catches it).
class FatalError : public std::runtime_error {
 public:
  explicit FatalError(const string& m) : std::runtime_error(m) {}
};

class LogMessage {
 public:
  LogMessage(const char* file, int line, bool fatal) : fatal_(fatal) { os_ << file << ":" << line << "] "; }
  ~LogMessage() noexcept(false) {
    if (fatal_) throw FatalError(os_.str());
    if (getenv("NPAIR_SHIM_VERBOSE")) fprintf(stderr, "%s\n", os_.str().c_str());
  }
  std::ostream& stream() { return os_; }
 private:
  std::ostringstream os_;
  bool fatal_;
};
struct LogVoidify { void operator&(std::ostream&) {} };

#define LOG_INFO ::caffe::LogMessage(__FILE__, __LINE__, false).stream()
#define LOG_WARNING ::caffe::LogMessage(__FILE__, __LINE__, false).stream()
#define LOG_ERROR ::caffe::LogMessage(__FILE__, __LINE__, false).stream()
#define LOG_FATAL ::caffe::LogMessage(__FILE__, __LINE__, true).stream()
#define LOG(severity) LOG_##severity
#define CHECK(cond) (cond) ? (void)0 : ::caffe::LogVoidify() & LOG_FATAL << "Check failed: " #cond " "
#define CHECK_OP(a, b, op) ((a)op(b)) ? (void)0 : ::caffe::LogVoidify() & LOG_FATAL << "Check failed: " #a " " #op " " #b " (" << (a) << " vs. " << (b) << ") "
#define CHECK_EQ(a, b) CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) CHECK_OP(a, b, !=)
#define CHECK_LE(a, b) CHECK_OP(a, b, <=)
#define CHECK_LT(a, b) CHECK_OP(a, b, <)
#define CHECK_GE(a, b) CHECK_OP(a, b, >=)
#define CHECK_GT(a, b) CHECK_OP(a, b, >)
#define NOT_IMPLEMENTED LOG(FATAL) << "Not Implemented Yet"

#define CUDA_CHECK(condition)                                                                     \
  do {                                                                                            \
    cudaError_t error__ = (condition);                                                            \
    CHECK_EQ(error__, cudaSuccess) << " " << cudaGetErrorString(error__);                         \
  } while (0)

#define INSTANTIATE_CLASS(classname) \
  template class classname<float>;   \
  template class classname<double>

// all device work of this layer sits behind the C ABI, so there are no per-layer __global__ functions to instantiate
#define INSTANTIATE_LAYER_GPU_FUNCS(classname)

class Caffe {
 public:
  enum Brew { CPU, GPU };
  static Brew mode() { return mode_; }
  static void set_mode(Brew m) { mode_ = m; }
  // ---- statics of the reference's private MPI fork ----
  static bool MULTI_GPU;
  static int NUM_GPU;
  static int RANK;
  // ---- B200 build: how a rank obtains the NCCL bootstrap id (the fork would MPI_Bcast it) ----
  static void set_nccl_unique_id(const void* id128);
  static const void* nccl_unique_id();   // NULL if none was set
 private:
  static Brew mode_;
};

}  // namespace caffe
#endif
