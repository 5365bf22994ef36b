// Mini-Caffe shim: Blob with Caffe's lazy host/device mirror (SyncedMemory head states) -- the part of
// caffe/blob.hpp + caffe/syncedmem.hpp the layer uses.  Host memory is pinned (Caffe does the same in GPU mode).
#ifndef CAFFE_BLOB_HPP_
#define CAFFE_BLOB_HPP_

#include <cuda_runtime.h>

#include "caffe/common.hpp"

namespace caffe {

class SyncedMemory {
 public:
  enum SyncedHead { UNINITIALIZED, HEAD_AT_CPU, HEAD_AT_GPU, SYNCED };
  explicit SyncedMemory(size_t size) : size_(size) {}
  ~SyncedMemory();
  const void* cpu_data();
  const void* gpu_data();
  void* mutable_cpu_data();
  void* mutable_gpu_data();
  // caffe/syncedmem.hpp: host -> device copy of a CPU-dirty blob on `stream`; the caller synchronises before use
  void async_gpu_push(const cudaStream_t& stream);
  SyncedHead head() const { return head_; }
  size_t size() const { return size_; }
 private:
  void to_cpu();
  void to_gpu();
  void* cpu_ptr_ = nullptr;
  void* gpu_ptr_ = nullptr;
  size_t size_;
  SyncedHead head_ = UNINITIALIZED;
};

template <typename Dtype>
class Blob {
 public:
  Blob() {}
  explicit Blob(const vector<int>& shape) { Reshape(shape); }
  Blob(int num, int channels, int height, int width) { Reshape(num, channels, height, width); }
  void Reshape(int num, int channels, int height, int width) {
    vector<int> s(4); s[0] = num; s[1] = channels; s[2] = height; s[3] = width; Reshape(s);
  }
  void Reshape(const vector<int>& shape) {
    int c = 1;
    for (size_t i = 0; i < shape.size(); ++i) { CHECK_GE(shape[i], 0); c *= shape[i]; }
    shape_ = shape; count_ = c;
    if (count_ > capacity_) {
      capacity_ = count_;
      data_.reset(new SyncedMemory(capacity_ * sizeof(Dtype)));
      diff_.reset(new SyncedMemory(capacity_ * sizeof(Dtype)));
    }
  }
  const vector<int>& shape() const { return shape_; }
  int shape(int i) const { return shape_[i]; }
  int num_axes() const { return static_cast<int>(shape_.size()); }
  int count() const { return count_; }
  int LegacyShape(int i) const { CHECK_LE(num_axes(), 4); return i < num_axes() ? shape_[i] : 1; }
  int num() const { return LegacyShape(0); }
  int channels() const { return LegacyShape(1); }
  int height() const { return LegacyShape(2); }
  int width() const { return LegacyShape(3); }
  const Dtype* cpu_data() const { CHECK(data_); return static_cast<const Dtype*>(data_->cpu_data()); }
  const Dtype* gpu_data() const { CHECK(data_); return static_cast<const Dtype*>(data_->gpu_data()); }
  const Dtype* cpu_diff() const { CHECK(diff_); return static_cast<const Dtype*>(diff_->cpu_data()); }
  const Dtype* gpu_diff() const { CHECK(diff_); return static_cast<const Dtype*>(diff_->gpu_data()); }
  Dtype* mutable_cpu_data() { CHECK(data_); return static_cast<Dtype*>(data_->mutable_cpu_data()); }
  Dtype* mutable_gpu_data() { CHECK(data_); return static_cast<Dtype*>(data_->mutable_gpu_data()); }
  Dtype* mutable_cpu_diff() { CHECK(diff_); return static_cast<Dtype*>(diff_->mutable_cpu_data()); }
  Dtype* mutable_gpu_diff() { CHECK(diff_); return static_cast<Dtype*>(diff_->mutable_gpu_data()); }
  const shared_ptr<SyncedMemory>& data() const { CHECK(data_); return data_; }
 private:
  shared_ptr<SyncedMemory> data_, diff_;
  vector<int> shape_;
  int count_ = 0, capacity_ = 0;
};

}  // namespace caffe
#endif
