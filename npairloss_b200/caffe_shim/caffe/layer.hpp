// Mini-Caffe shim: Layer<Dtype> with the SetUp / Forward / Backward wrappers of BVLC caffe/layer.hpp.
#ifndef CAFFE_LAYER_HPP_
#define CAFFE_LAYER_HPP_

#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/proto/caffe.pb.h"

namespace caffe {

template <typename Dtype>
class Layer {
 public:
  explicit Layer(const LayerParameter& param) : layer_param_(param) {}
  virtual ~Layer() {}
  void SetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    CheckBlobCounts(bottom, top);
    LayerSetUp(bottom, top);
    Reshape(bottom, top);
    SetLossWeights(top);
  }
  virtual void LayerSetUp(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) = 0;
  // returns the weighted loss like Layer::Forward (tops of this layer are host scalars)
  Dtype Forward(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    Reshape(bottom, top);
    if (Caffe::mode() == Caffe::GPU) Forward_gpu(bottom, top); else Forward_cpu(bottom, top);
    Dtype loss = 0;
    for (size_t t = 0; t < top.size(); ++t) {
      if (!this->loss(static_cast<int>(t))) continue;
      const Dtype* d = top[t]->cpu_data();
      const Dtype* w = top[t]->cpu_diff();
      for (int i = 0; i < top[t]->count(); ++i) loss += d[i] * w[i];
    }
    return loss;
  }
  void Backward(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, const vector<Blob<Dtype>*>& bottom) {
    if (Caffe::mode() == Caffe::GPU) Backward_gpu(top, propagate_down, bottom); else Backward_cpu(top, propagate_down, bottom);
  }
  const LayerParameter& layer_param() const { return layer_param_; }
  virtual inline const char* type() const { return ""; }
  virtual inline int ExactNumBottomBlobs() const { return -1; }
  virtual inline int MinBottomBlobs() const { return -1; }
  virtual inline int MaxBottomBlobs() const { return -1; }
  virtual inline int ExactNumTopBlobs() const { return -1; }
  virtual inline int MinTopBlobs() const { return -1; }
  virtual inline int MaxTopBlobs() const { return -1; }
  inline Dtype loss(int top_index) const { return (static_cast<int>(loss_.size()) > top_index) ? loss_[top_index] : Dtype(0); }
 protected:
  LayerParameter layer_param_;
  vector<Dtype> loss_;
  virtual void Forward_cpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) = 0;
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) { Forward_cpu(bottom, top); }
  virtual void Backward_cpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, const vector<Blob<Dtype>*>& bottom) = 0;
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, const vector<Blob<Dtype>*>& bottom) {
    Backward_cpu(top, propagate_down, bottom);
  }
  virtual void CheckBlobCounts(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    if (ExactNumBottomBlobs() >= 0) CHECK_EQ(ExactNumBottomBlobs(), static_cast<int>(bottom.size())) << type() << " Layer takes " << ExactNumBottomBlobs() << " bottom blob(s) as input.";
    if (MinBottomBlobs() >= 0) CHECK_LE(MinBottomBlobs(), static_cast<int>(bottom.size()));
    if (MaxBottomBlobs() >= 0) CHECK_GE(MaxBottomBlobs(), static_cast<int>(bottom.size()));
    if (ExactNumTopBlobs() >= 0) CHECK_EQ(ExactNumTopBlobs(), static_cast<int>(top.size()));
    if (MinTopBlobs() >= 0) CHECK_LE(MinTopBlobs(), static_cast<int>(top.size())) << type() << " Layer produces at least " << MinTopBlobs() << " top blob(s) as output.";
    if (MaxTopBlobs() >= 0) CHECK_GE(MaxTopBlobs(), static_cast<int>(top.size())) << type() << " Layer produces at most " << MaxTopBlobs() << " top blob(s) as output.";
  }
  inline void SetLossWeights(const vector<Blob<Dtype>*>& top) {
    const int n = layer_param_.loss_weight_size();
    if (!n) return;
    CHECK_EQ(static_cast<int>(top.size()), n) << "loss_weight must be unspecified or specified once per top blob.";
    for (int t = 0; t < n; ++t) {
      const Dtype w = layer_param_.loss_weight(t);
      if (w == Dtype(0)) continue;
      if (static_cast<int>(loss_.size()) <= t) loss_.resize(t + 1, Dtype(0));
      loss_[t] = w;
      Dtype* d = top[t]->mutable_cpu_diff();
      for (int i = 0; i < top[t]->count(); ++i) d[i] = w;
    }
  }
};

}  // namespace caffe
#endif
