// L2NormalizeLayer -- the producer layer in front of the loss in the reference net (usage/def.prototxt:115-120, type "L2Normalize").
// Its source is not part of the reference tree (a layer of the private Caffe fork); semantics as stated in include/npair_b200.h:
// y = x / ||x||_2 per sample, backward dx = (dy - y (y . dy)) / ||x||.  Host class only: the kernels sit behind the C ABI
// (npair_l2normalize_forward / _backward).  float only (the device path is fp32).
#ifndef CAFFE_L2_NORMALIZE_LAYER_HPP_
#define CAFFE_L2_NORMALIZE_LAYER_HPP_

#include <vector>

#include "caffe/blob.hpp"
#include "caffe/layer.hpp"

namespace caffe {

template <typename Dtype>
class L2NormalizeLayer : public Layer<Dtype> {
 public:
  explicit L2NormalizeLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual ~L2NormalizeLayer();
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "L2Normalize"; }
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }

 protected:
  virtual void Forward_cpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Backward_cpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, const vector<Blob<Dtype>*>& bottom);
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, const vector<Blob<Dtype>*>& bottom);

 private:
  float* inv_norm_ = nullptr;    // device, num floats
  int num_ = 0, dim_ = 0;
};

}  // namespace caffe
#endif
