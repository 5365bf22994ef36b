// Synthetic training harness: runs the reference's layer chain for many iterations inside the shim, the way
// `caffe train --solver usage/solver.prototxt` drives it:
//     MultibatchData (usage/def.prototxt:2-59)  ->  [trunk, elided in the reference: usage/def.prototxt:112-114]  ->  L2Normalize (:115-120)
//     ->  NPairMultiClassLoss (:121-151),   SGD with momentum / weight decay / "step" learning-rate policy (usage/solver.prototxt:1-17).
// The reference ships neither the data layer nor the trunk; both are SYNTHETIC stand-ins here (host code, not product):
//   * MultibatchDataLayer keeps the one property the loss relies on -- every batch holds identity_num_per_batch identities x
//     img_num_per_identity images, images of an identity contiguous (labels as floats) -- over an in-memory set of sample ids;
//   * SyntheticTrunkLayer is a learnable embedding table (sample id -> D floats) initialised as class centre + noise, so the loss
//     has something to train.  It lives on the host: its top is pushed to the device by Blob like any data layer's output, and it reads
//     the gradient back through cpu_diff().
// Everything on the GPU (L2Normalize, the loss layer) goes through the registered layer classes and the C ABI.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/layer_factory.hpp"
#include "caffe/proto/caffe.pb.h"

using namespace caffe;

namespace {

int extra_int(const LayerParameter& p, const std::string& key, int dflt) {
  const std::string* v = p.extra(key);
  return v ? atoi(v->c_str()) : dflt;
}

// ---- MultibatchData stand-in: tops = {sample ids (num x 1 x 1 x 1, as floats), labels (num)} ----
template <typename Dtype>
class MultibatchDataLayer : public Layer<Dtype> {
 public:
  explicit MultibatchDataLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual inline const char* type() const { return "MultibatchData"; }
  virtual inline int ExactNumBottomBlobs() const { return 0; }
  virtual inline int ExactNumTopBlobs() const { return 2; }
  void Configure(int num_identities, int imgs_total, unsigned seed) { n_id_ = num_identities; imgs_total_ = imgs_total; rng_.seed(seed); }
  virtual void LayerSetUp(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) {
    const LayerParameter& p = this->layer_param();
    ids_per_batch_ = extra_int(p, "multi_batch_data_param.identity_num_per_batch", 60);      // usage/def.prototxt:25
    imgs_per_id_ = extra_int(p, "multi_batch_data_param.img_num_per_identity", 2);           // :26
    const int bs = extra_int(p, "multi_batch_data_param.batch_size", ids_per_batch_ * imgs_per_id_);
    CHECK_EQ(bs, ids_per_batch_ * imgs_per_id_) << "batch_size must equal identity_num_per_batch * img_num_per_identity";
    CHECK_GE(n_id_, ids_per_batch_) << "synthetic set has fewer identities than one batch needs";
    CHECK_GE(imgs_total_, imgs_per_id_);
    order_.resize(n_id_);
    for (int i = 0; i < n_id_; ++i) order_[i] = i;
  }
  virtual void Reshape(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>& top) {
    const int n = ids_per_batch_ * imgs_per_id_;
    top[0]->Reshape(n, 1, 1, 1);
    vector<int> s(1, n);
    top[1]->Reshape(s);
  }
  int batch() const { return ids_per_batch_ * imgs_per_id_; }
  int imgs_total() const { return imgs_total_; }
 protected:
  virtual void Forward_cpu(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>& top) {
    Dtype* ids = top[0]->mutable_cpu_data();
    Dtype* lab = top[1]->mutable_cpu_data();
    for (int k = 0; k < ids_per_batch_; ++k) {                 // rand_identity: true (usage/def.prototxt:27): partial Fisher-Yates
      std::uniform_int_distribution<int> d(k, n_id_ - 1);
      std::swap(order_[k], order_[d(rng_)]);
      const int id = order_[k];
      std::vector<int> imgs(imgs_total_);
      for (int m = 0; m < imgs_total_; ++m) imgs[m] = m;
      for (int m = 0; m < imgs_per_id_; ++m) {
        std::uniform_int_distribution<int> e(m, imgs_total_ - 1);
        std::swap(imgs[m], imgs[e(rng_)]);
        ids[k * imgs_per_id_ + m] = static_cast<Dtype>(id * imgs_total_ + imgs[m]);
        lab[k * imgs_per_id_ + m] = static_cast<Dtype>(id);
      }
    }
  }
  virtual void Backward_cpu(const vector<Blob<Dtype>*>&, const vector<bool>&, const vector<Blob<Dtype>*>&) {}
 private:
  int n_id_ = 0, imgs_total_ = 0, ids_per_batch_ = 0, imgs_per_id_ = 0;
  std::vector<int> order_;
  std::mt19937 rng_;
};

// ---- the elided trunk: a learnable embedding table on the host ----
template <typename Dtype>
class SyntheticTrunkLayer : public Layer<Dtype> {
 public:
  explicit SyntheticTrunkLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual inline const char* type() const { return "SyntheticTrunk"; }
  void Configure(int n_samples, int imgs_total, int dim, float noise, unsigned seed) {
    dim_ = dim;
    W_.assign(static_cast<size_t>(n_samples) * dim, 0.f); V_.assign(W_.size(), 0.f);
    std::mt19937 g(seed); std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> centre(dim);
    for (int s = 0; s < n_samples; ++s) {
      if (s % imgs_total == 0) for (int d = 0; d < dim; ++d) centre[d] = nd(g);
      for (int d = 0; d < dim; ++d) W_[static_cast<size_t>(s) * dim + d] = centre[d] + noise * nd(g);
    }
  }
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) { top[0]->Reshape(bottom[0]->num(), dim_, 1, 1); }
  // SGD with momentum and weight decay on the rows of the last batch (Caffe's SGDSolver::ComputeUpdateValue on a sparse parameter)
  void Update(float lr, float momentum, float weight_decay) {
    for (size_t k = 0; k < rows_.size(); ++k) {
      float* w = &W_[static_cast<size_t>(rows_[k]) * dim_];
      float* v = &V_[static_cast<size_t>(rows_[k]) * dim_];
      const float* g = &grad_[k * dim_];
      for (int d = 0; d < dim_; ++d) { v[d] = momentum * v[d] + lr * (g[d] + weight_decay * w[d]); w[d] -= v[d]; }
    }
  }
 protected:
  virtual void Forward_cpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    const int n = bottom[0]->num();
    const Dtype* ids = bottom[0]->cpu_data();
    Dtype* out = top[0]->mutable_cpu_data();                     // CPU-dirty: the next layer's gpu_data() pays the H2D copy
    rows_.resize(n);
    for (int r = 0; r < n; ++r) {
      rows_[r] = static_cast<int>(ids[r]);
      const float* w = &W_[static_cast<size_t>(rows_[r]) * dim_];
      for (int d = 0; d < dim_; ++d) out[static_cast<size_t>(r) * dim_ + d] = static_cast<Dtype>(w[d]);
    }
  }
  virtual void Backward_cpu(const vector<Blob<Dtype>*>& top, const vector<bool>&, const vector<Blob<Dtype>*>&) {
    const Dtype* g = top[0]->cpu_diff();                         // D2H of the gradient the L2Normalize layer wrote on the device
    grad_.assign(g, g + static_cast<size_t>(top[0]->num()) * dim_);
  }
  // the trunk runs on the host even in GPU mode
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& b, const vector<Blob<Dtype>*>& t) { Forward_cpu(b, t); }
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& t, const vector<bool>& p, const vector<Blob<Dtype>*>& b) { Backward_cpu(t, p, b); }
 private:
  int dim_ = 0;
  std::vector<float> W_, V_, grad_;
  std::vector<int> rows_;
};

REGISTER_LAYER_CLASS(MultibatchData);
REGISTER_LAYER_CLASS(SyntheticTrunk);

thread_local std::string g_serr;
}  // namespace

extern "C" {

const char* npc_solver_last_error(void) { return g_serr.c_str(); }

// Runs `iters` (or the solver's max_iter when iters <= 0) SGD iterations of the TRAIN-phase chain found in `net_prototxt`.
// log: rows of {iter, weighted loss as Caffe prints it, top[0..4]} every `display` iterations (solver prototxt) and at the last one.
// Returns the number of rows written, or -1 (npc_solver_last_error()).
int npc_solver_run(const char* net_prototxt, const char* solver_prototxt, int feature_dim, int num_identities, int imgs_total, int iters,
                   unsigned seed, float noise, float* log, int max_rows) {
  try {
    std::vector<LayerParameter> all;
    std::string perr;
    if (!ReadLayersFromText(net_prototxt ? net_prototxt : "", &all, &perr)) { g_serr = "net prototxt: " + perr; return -1; }
    std::map<std::string, std::string> sp;
    if (!ReadScalarsFromText(solver_prototxt ? solver_prototxt : "", &sp, &perr)) { g_serr = "solver prototxt: " + perr; return -1; }
    auto sval = [&](const char* k, double d) { return sp.count(k) ? atof(sp[k].c_str()) : d; };
    const float base_lr = static_cast<float>(sval("base_lr", 0.001)), gamma = static_cast<float>(sval("gamma", 0.5));
    const float momentum = static_cast<float>(sval("momentum", 0.9)), wd = static_cast<float>(sval("weight_decay", 0.00002));
    const int stepsize = static_cast<int>(sval("stepsize", 10000)), display = std::max(1, static_cast<int>(sval("display", 100)));
    const std::string policy = sp.count("lr_policy") ? sp["lr_policy"] : "fixed";
    const int max_iter = iters > 0 ? iters : static_cast<int>(sval("max_iter", 1000));
    // TRAIN-phase layers with a registered type, in file order
    const LayerParameter *pd = nullptr, *pn = nullptr, *pl = nullptr;
    for (size_t i = 0; i < all.size(); ++i) {
      const std::string* ph = all[i].extra("include.phase");
      if (ph && *ph != "TRAIN") continue;
      if (all[i].type() == "MultibatchData" && !pd) pd = &all[i];
      else if (all[i].type() == "L2Normalize" && !pn) pn = &all[i];
      else if (all[i].type() == "NPairMultiClassLoss" && !pl) pl = &all[i];
    }
    if (!pd || !pl) { g_serr = "the net needs a TRAIN-phase MultibatchData layer and an NPairMultiClassLoss layer"; return -1; }
    Caffe::set_mode(Caffe::GPU);
    Caffe::NUM_GPU = 1; Caffe::RANK = 0; Caffe::MULTI_GPU = false;
    // blobs
    Blob<float> ids, labels, feat, feat_norm;
    std::vector<Blob<float> > tops(pl->top_size());
    shared_ptr<Layer<float> > data = LayerRegistry<float>::CreateLayer(*pd);
    MultibatchDataLayer<float>* dl = static_cast<MultibatchDataLayer<float>*>(data.get());
    dl->Configure(num_identities, imgs_total, seed);
    std::vector<Blob<float>*> d_bot, d_top; d_top.push_back(&ids); d_top.push_back(&labels);
    data->SetUp(d_bot, d_top);
    LayerParameter tp; tp.set_name("synthetic_trunk"); tp.set_type("SyntheticTrunk");
    shared_ptr<Layer<float> > trunk = LayerRegistry<float>::CreateLayer(tp);
    SyntheticTrunkLayer<float>* tl = static_cast<SyntheticTrunkLayer<float>*>(trunk.get());
    tl->Configure(num_identities * imgs_total, imgs_total, feature_dim, noise, seed + 1);
    std::vector<Blob<float>*> t_bot(1, &ids), t_top(1, &feat);
    trunk->SetUp(t_bot, t_top);
    shared_ptr<Layer<float> > norm;
    std::vector<Blob<float>*> n_bot(1, &feat), n_top(1, &feat_norm);
    if (pn) { norm = LayerRegistry<float>::CreateLayer(*pn); norm->SetUp(n_bot, n_top); }
    shared_ptr<Layer<float> > loss = LayerRegistry<float>::CreateLayer(*pl);
    std::vector<Blob<float>*> l_bot, l_top;
    l_bot.push_back(pn ? &feat_norm : &feat); l_bot.push_back(&labels);
    for (size_t t = 0; t < tops.size(); ++t) l_top.push_back(&tops[t]);
    loss->SetUp(l_bot, l_top);
    std::vector<bool> pd_all(2, false); pd_all[0] = true;
    int rows = 0;
    for (int it = 0; it < max_iter; ++it) {
      data->Forward(d_bot, d_top);
      trunk->Forward(t_bot, t_top);
      if (norm) norm->Forward(n_bot, n_top);
      const float wl = loss->Forward(l_bot, l_top);
      loss->Backward(l_top, pd_all, l_bot);
      if (norm) norm->Backward(n_top, std::vector<bool>(1, true), n_bot);
      trunk->Backward(t_top, std::vector<bool>(1, false), t_bot);
      float lr = base_lr;
      if (policy == "step") lr = base_lr * std::pow(gamma, static_cast<float>(it / std::max(1, stepsize)));     // usage/solver.prototxt:8-10
      tl->Update(lr, momentum, wd);
      if ((it % display == 0 || it == max_iter - 1) && log && rows < max_rows) {
        float* r = log + 7 * rows++;
        r[0] = static_cast<float>(it); r[1] = wl;
        for (int t = 0; t < 5; ++t) r[2 + t] = t < static_cast<int>(tops.size()) ? tops[t].cpu_data()[0] : 0.f;
      }
    }
    return rows;
  } catch (const std::exception& e) { g_serr = e.what(); return -1; }
}

}  // extern "C"
