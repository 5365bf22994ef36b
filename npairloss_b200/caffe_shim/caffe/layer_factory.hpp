// Mini-Caffe shim: LayerRegistry + REGISTER_LAYER_CLASS (BVLC caffe/layer_factory.hpp).
#ifndef CAFFE_LAYER_FACTORY_HPP_
#define CAFFE_LAYER_FACTORY_HPP_
#include <map>
#include "caffe/layer.hpp"
namespace caffe {
template <typename Dtype>
class LayerRegistry {
 public:
  typedef shared_ptr<Layer<Dtype> > (*Creator)(const LayerParameter&);
  typedef std::map<string, Creator> CreatorRegistry;
  static CreatorRegistry& Registry() { static CreatorRegistry* g = new CreatorRegistry(); return *g; }
  static void AddCreator(const string& type, Creator c) { Registry()[type] = c; }
  static shared_ptr<Layer<Dtype> > CreateLayer(const LayerParameter& param) {
    typename CreatorRegistry::iterator it = Registry().find(param.type());
    CHECK(it != Registry().end()) << "Unknown layer type: " << param.type();
    return it->second(param);
  }
};
template <typename Dtype>
class LayerRegisterer {
 public:
  LayerRegisterer(const string& type, shared_ptr<Layer<Dtype> > (*creator)(const LayerParameter&)) { LayerRegistry<Dtype>::AddCreator(type, creator); }
};
#define REGISTER_LAYER_CREATOR(type, creator)                                  \
  static LayerRegisterer<float> g_creator_f_##type(#type, creator<float>);     \
  static LayerRegisterer<double> g_creator_d_##type(#type, creator<double>)
#define REGISTER_LAYER_CLASS(type)                                                               \
  template <typename Dtype>                                                                      \
  shared_ptr<Layer<Dtype> > Creator_##type##Layer(const LayerParameter& param) {                 \
    return shared_ptr<Layer<Dtype> >(new type##Layer<Dtype>(param));                             \
  }                                                                                              \
  REGISTER_LAYER_CREATOR(type, Creator_##type##Layer)
}  // namespace caffe
#endif
