// L2NormalizeLayer host side (see the header).  No CPU path, like the loss layer.
#include "l2_normalize_layer.hpp"

#include <cuda_runtime.h>

#include "caffe/layer_factory.hpp"
#include "npair_b200.h"

namespace caffe {

template <typename Dtype>
L2NormalizeLayer<Dtype>::~L2NormalizeLayer() { if (inv_norm_) cudaFree(inv_norm_); }

template <typename Dtype>
void L2NormalizeLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>&) {
  CHECK(sizeof(Dtype) == 4) << "L2Normalize: the B200 build is float only";
  num_ = bottom[0]->num();
  dim_ = bottom[0]->count() / (num_ > 0 ? num_ : 1);
  if (inv_norm_) { cudaFree(inv_norm_); inv_norm_ = nullptr; }
  CUDA_CHECK(cudaMalloc(&inv_norm_, sizeof(float) * (num_ > 0 ? num_ : 1)));
}

template <typename Dtype>
void L2NormalizeLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_EQ(bottom[0]->num(), num_) << "L2Normalize: batch size changed after LayerSetUp";
  top[0]->Reshape(bottom[0]->shape());
}

template <typename Dtype>
void L2NormalizeLayer<Dtype>::Forward_cpu(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) {
  LOG(FATAL) << "L2NormalizeLayer has no CPU path in this build; use Caffe::GPU on a B200";
}
template <typename Dtype>
void L2NormalizeLayer<Dtype>::Backward_cpu(const vector<Blob<Dtype>*>&, const vector<bool>&, const vector<Blob<Dtype>*>&) {
  LOG(FATAL) << "L2NormalizeLayer has no CPU path in this build; use Caffe::GPU on a B200";
}

template <typename Dtype>
void L2NormalizeLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const int rc = npair_l2normalize_forward(reinterpret_cast<const float*>(bottom[0]->gpu_data()), num_, dim_,
                                           reinterpret_cast<float*>(top[0]->mutable_gpu_data()), inv_norm_, nullptr);
  CHECK_EQ(rc, NPAIR_OK) << "npair_l2normalize_forward: " << npair_last_error(nullptr);
}

template <typename Dtype>
void L2NormalizeLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, const vector<Blob<Dtype>*>& bottom) {
  if (!propagate_down.empty() && !propagate_down[0]) return;
  const int rc = npair_l2normalize_backward(reinterpret_cast<const float*>(top[0]->gpu_data()), inv_norm_, reinterpret_cast<const float*>(top[0]->gpu_diff()),
                                            num_, dim_, reinterpret_cast<float*>(bottom[0]->mutable_gpu_diff()), nullptr);
  CHECK_EQ(rc, NPAIR_OK) << "npair_l2normalize_backward: " << npair_last_error(nullptr);
}

INSTANTIATE_CLASS(L2NormalizeLayer);
REGISTER_LAYER_CLASS(L2Normalize);

}  // namespace caffe
