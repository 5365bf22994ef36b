// Mini-Caffe shim: LossLayer base (BVLC caffe/loss_layer.hpp; the fork's header is called loss_layers.hpp,
// npair_multi_class_loss.cpp:1).
#ifndef CAFFE_LOSS_LAYERS_HPP_
#define CAFFE_LOSS_LAYERS_HPP_
#include "caffe/layer.hpp"
namespace caffe {
template <typename Dtype>
class LossLayer : public Layer<Dtype> {
 public:
  explicit LossLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual inline int ExactNumBottomBlobs() const { return 2; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
};
}  // namespace caffe
#endif
