// Mini-Caffe shim: hand-written stand-in for the protoc output of caffe.proto, restricted to LayerParameter fields the
// layer block of usage/def.prototxt:121-151 uses plus the reference's NPairLossParameter (caffe.proto:2-23: field
// npair_loss_param = 8866720, four floats, two enums, defaults 0, 0, -1, -1, LOCAL, RAND, LOCAL, RAND).
// protoc is not available in this environment; with a real Caffe tree paste reference caffe.proto:2-23 into
// src/caffe/proto/caffe.proto instead.  Accessor names match what protoc generates.
#ifndef CAFFE_PROTO_CAFFE_PB_H_
#define CAFFE_PROTO_CAFFE_PB_H_

#include <map>
#include <string>
#include <vector>

namespace caffe {

enum NPairLossParameter_MiningRegion { NPairLossParameter_MiningRegion_GLOBAL = 0, NPairLossParameter_MiningRegion_LOCAL = 1 };
enum NPairLossParameter_MiningMethod {
  NPairLossParameter_MiningMethod_HARD = 0,
  NPairLossParameter_MiningMethod_EASY = 1,
  NPairLossParameter_MiningMethod_RAND = 2,
  NPairLossParameter_MiningMethod_RELATIVE_HARD = 3,
  NPairLossParameter_MiningMethod_RELATIVE_EASY = 4
};

class NPairLossParameter {
 public:
  typedef NPairLossParameter_MiningRegion MiningRegion;
  typedef NPairLossParameter_MiningMethod MiningMethod;
  static const MiningRegion GLOBAL = NPairLossParameter_MiningRegion_GLOBAL;
  static const MiningRegion LOCAL = NPairLossParameter_MiningRegion_LOCAL;
  static const MiningMethod HARD = NPairLossParameter_MiningMethod_HARD;
  static const MiningMethod EASY = NPairLossParameter_MiningMethod_EASY;
  static const MiningMethod RAND = NPairLossParameter_MiningMethod_RAND;
  static const MiningMethod RELATIVE_HARD = NPairLossParameter_MiningMethod_RELATIVE_HARD;
  static const MiningMethod RELATIVE_EASY = NPairLossParameter_MiningMethod_RELATIVE_EASY;

  float margin_ident() const { return margin_ident_; }            // field 1, default 0
  float margin_diff() const { return margin_diff_; }              // field 8, default 0
  float identsn() const { return identsn_; }                      // field 2, default -1
  float diffsn() const { return diffsn_; }                        // field 3, default -1
  MiningRegion ap_mining_region() const { return ap_region_; }    // field 4, default LOCAL
  MiningMethod ap_mining_method() const { return ap_method_; }    // field 5, default RAND
  MiningRegion an_mining_region() const { return an_region_; }    // field 6, default LOCAL
  MiningMethod an_mining_method() const { return an_method_; }    // field 7, default RAND
  void set_margin_ident(float v) { margin_ident_ = v; }
  void set_margin_diff(float v) { margin_diff_ = v; }
  void set_identsn(float v) { identsn_ = v; }
  void set_diffsn(float v) { diffsn_ = v; }
  void set_ap_mining_region(MiningRegion v) { ap_region_ = v; }
  void set_ap_mining_method(MiningMethod v) { ap_method_ = v; }
  void set_an_mining_region(MiningRegion v) { an_region_ = v; }
  void set_an_mining_method(MiningMethod v) { an_method_ = v; }
 private:
  float margin_ident_ = 0.f, margin_diff_ = 0.f, identsn_ = -1.f, diffsn_ = -1.f;
  MiningRegion ap_region_ = NPairLossParameter_MiningRegion_LOCAL, an_region_ = NPairLossParameter_MiningRegion_LOCAL;
  MiningMethod ap_method_ = NPairLossParameter_MiningMethod_RAND, an_method_ = NPairLossParameter_MiningMethod_RAND;
};

class LayerParameter {
 public:
  const std::string& name() const { return name_; }
  const std::string& type() const { return type_; }
  int bottom_size() const { return static_cast<int>(bottom_.size()); }
  int top_size() const { return static_cast<int>(top_.size()); }
  const std::string& bottom(int i) const { return bottom_[i]; }
  const std::string& top(int i) const { return top_[i]; }
  int loss_weight_size() const { return static_cast<int>(loss_weight_.size()); }
  float loss_weight(int i) const { return loss_weight_[i]; }
  bool has_npair_loss_param() const { return has_npair_; }
  const NPairLossParameter& npair_loss_param() const { return npair_; }
  NPairLossParameter* mutable_npair_loss_param() { has_npair_ = true; return &npair_; }
  void set_name(const std::string& s) { name_ = s; }
  void set_type(const std::string& s) { type_ = s; }
  void add_bottom(const std::string& s) { bottom_.push_back(s); }
  void add_top(const std::string& s) { top_.push_back(s); }
  void add_loss_weight(float w) { loss_weight_.push_back(w); }
  // scalar fields of the layer's OTHER nested messages, kept as text under "message.field" (first occurrence), e.g.
  // "multi_batch_data_param.identity_num_per_batch" or "include.phase": enough for the synthetic data layer of the harness
  const std::string* extra(const std::string& key) const { std::map<std::string, std::string>::const_iterator it = extra_.find(key); return it == extra_.end() ? 0 : &it->second; }
  void set_extra(const std::string& key, const std::string& v) { if (!extra_.count(key)) extra_[key] = v; }
 private:
  std::map<std::string, std::string> extra_;
  std::string name_, type_;
  std::vector<std::string> bottom_, top_;
  std::vector<float> loss_weight_;
  NPairLossParameter npair_;
  bool has_npair_ = false;
};

// Minimal protobuf text-format reader: returns every `layer { ... }` / `layers { ... }` block of a prototxt.  Unknown
// fields and nested messages are skipped; `#` comments and the literal "." placeholder lines of usage/def.prototxt are
// ignored.  Unknown enum identifiers inside npair_loss_param are an error.
bool ReadLayersFromText(const std::string& text, std::vector<LayerParameter>* layers, std::string* error);
// top-level "key: value" pairs (solver prototxt)
bool ReadScalarsFromText(const std::string& text, std::map<std::string, std::string>* out, std::string* error);

}  // namespace caffe
#endif
