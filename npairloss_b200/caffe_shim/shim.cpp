// Mini-Caffe shim implementation: SyncedMemory, Caffe statics and the prototxt reader.
#include <cuda_runtime.h>

#include <cctype>
#include <cstring>
#include <map>

#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/proto/caffe.pb.h"

namespace caffe {

Caffe::Brew Caffe::mode_ = Caffe::GPU;
bool Caffe::MULTI_GPU = false;
int Caffe::NUM_GPU = 1;
int Caffe::RANK = 0;
static unsigned char g_nccl_id[128];
static bool g_have_nccl_id = false;
void Caffe::set_nccl_unique_id(const void* id128) {
  if (id128) { memcpy(g_nccl_id, id128, 128); g_have_nccl_id = true; } else g_have_nccl_id = false;
}
const void* Caffe::nccl_unique_id() { return g_have_nccl_id ? g_nccl_id : nullptr; }

// ---------------------------------------------------------------- SyncedMemory (caffe/syncedmem.cpp semantics)
SyncedMemory::~SyncedMemory() {
  if (cpu_ptr_) cudaFreeHost(cpu_ptr_);
  if (gpu_ptr_) cudaFree(gpu_ptr_);
}
void SyncedMemory::to_cpu() {
  switch (head_) {
    case UNINITIALIZED:
      CUDA_CHECK(cudaMallocHost(&cpu_ptr_, size_ ? size_ : 1));
      memset(cpu_ptr_, 0, size_);
      head_ = HEAD_AT_CPU;
      break;
    case HEAD_AT_GPU:
      if (!cpu_ptr_) CUDA_CHECK(cudaMallocHost(&cpu_ptr_, size_ ? size_ : 1));
      CUDA_CHECK(cudaMemcpy(cpu_ptr_, gpu_ptr_, size_, cudaMemcpyDeviceToHost));   // blocking, like caffe_gpu_memcpy
      head_ = SYNCED;
      break;
    default: break;
  }
}
void SyncedMemory::to_gpu() {
  switch (head_) {
    case UNINITIALIZED:
      CUDA_CHECK(cudaMalloc(&gpu_ptr_, size_ ? size_ : 1));
      CUDA_CHECK(cudaMemset(gpu_ptr_, 0, size_));
      head_ = HEAD_AT_GPU;
      break;
    case HEAD_AT_CPU:
      if (!gpu_ptr_) CUDA_CHECK(cudaMalloc(&gpu_ptr_, size_ ? size_ : 1));
      CUDA_CHECK(cudaMemcpy(gpu_ptr_, cpu_ptr_, size_, cudaMemcpyHostToDevice));
      head_ = SYNCED;
      break;
    default: break;
  }
}
void SyncedMemory::async_gpu_push(const cudaStream_t& stream) {
  CHECK(head_ == HEAD_AT_CPU) << "async_gpu_push needs a CPU-dirty blob";
  if (!gpu_ptr_) CUDA_CHECK(cudaMalloc(&gpu_ptr_, size_ ? size_ : 1));
  CUDA_CHECK(cudaMemcpyAsync(gpu_ptr_, cpu_ptr_, size_, cudaMemcpyHostToDevice, stream));
  head_ = SYNCED;                                  // the caller orders its consumers after `stream` (event / synchronize)
}
const void* SyncedMemory::cpu_data() { to_cpu(); return cpu_ptr_; }
const void* SyncedMemory::gpu_data() { to_gpu(); return gpu_ptr_; }
void* SyncedMemory::mutable_cpu_data() { to_cpu(); head_ = HEAD_AT_CPU; return cpu_ptr_; }
void* SyncedMemory::mutable_gpu_data() { to_gpu(); head_ = HEAD_AT_GPU; return gpu_ptr_; }

// ---------------------------------------------------------------- prototxt text reader
namespace {
struct Tok { enum Kind { IDENT, NUMBER, STRING, LBRACE, RBRACE, COLON, END } kind; std::string text; };

class Lexer {
 public:
  explicit Lexer(const std::string& s) : s_(s) {}
  Tok next() {
    for (;;) {
      while (i_ < s_.size() && (isspace(static_cast<unsigned char>(s_[i_])) || s_[i_] == ',' || s_[i_] == ';')) ++i_;
      if (i_ < s_.size() && s_[i_] == '#') { while (i_ < s_.size() && s_[i_] != '\n') ++i_; continue; }
      // a lone '.' (the elision marks of usage/def.prototxt:112-114) is not a token
      if (i_ < s_.size() && s_[i_] == '.' && (i_ + 1 >= s_.size() || !isdigit(static_cast<unsigned char>(s_[i_ + 1])))) { ++i_; continue; }
      break;
    }
    if (i_ >= s_.size()) return {Tok::END, ""};
    const char c = s_[i_];
    if (c == '{' || c == '<') { ++i_; return {Tok::LBRACE, "{"}; }
    if (c == '}' || c == '>') { ++i_; return {Tok::RBRACE, "}"}; }
    if (c == ':') { ++i_; return {Tok::COLON, ":"}; }
    if (c == '"' || c == '\'') {
      const char q = c; std::string v; ++i_;
      while (i_ < s_.size() && s_[i_] != q) { if (s_[i_] == '\\' && i_ + 1 < s_.size()) ++i_; v.push_back(s_[i_++]); }
      ++i_;
      return {Tok::STRING, v};
    }
    if (isdigit(static_cast<unsigned char>(c)) || c == '-' || c == '+' || c == '.') {
      size_t j = i_ + 1;
      while (j < s_.size() && (isalnum(static_cast<unsigned char>(s_[j])) || s_[j] == '.' || s_[j] == '-' || s_[j] == '+')) ++j;
      Tok t{Tok::NUMBER, s_.substr(i_, j - i_)}; i_ = j; return t;
    }
    if (isalpha(static_cast<unsigned char>(c)) || c == '_') {
      size_t j = i_ + 1;
      while (j < s_.size() && (isalnum(static_cast<unsigned char>(s_[j])) || s_[j] == '_')) ++j;
      Tok t{Tok::IDENT, s_.substr(i_, j - i_)}; i_ = j; return t;
    }
    ++i_;
    return next();
  }
 private:
  const std::string& s_;
  size_t i_ = 0;
};

bool skip_message(Lexer& lx, std::string* err) {
  int depth = 1;
  for (;;) {
    Tok t = lx.next();
    if (t.kind == Tok::END) { *err = "unterminated message"; return false; }
    if (t.kind == Tok::LBRACE) ++depth;
    if (t.kind == Tok::RBRACE && --depth == 0) return true;
  }
}

bool parse_float(const Tok& t, float* v) {
  if (t.kind != Tok::NUMBER) return false;
  char* end = nullptr;
  std::string s = t.text;
  if (!s.empty() && (s.back() == 'f' || s.back() == 'F')) s.pop_back();
  *v = strtof(s.c_str(), &end);
  return end && *end == 0;
}

bool parse_npair(Lexer& lx, NPairLossParameter* p, std::string* err) {
  for (;;) {
    Tok k = lx.next();
    if (k.kind == Tok::RBRACE) return true;
    if (k.kind != Tok::IDENT) { *err = "npair_loss_param: expected field name"; return false; }
    Tok c = lx.next();
    if (c.kind == Tok::LBRACE) { if (!skip_message(lx, err)) return false; continue; }
    if (c.kind != Tok::COLON) { *err = "npair_loss_param: expected ':' after " + k.text; return false; }
    Tok v = lx.next();
    float f = 0.f;
    auto region = [&](NPairLossParameter::MiningRegion* out) {
      if (v.text == "GLOBAL" || v.text == "0") { *out = NPairLossParameter::GLOBAL; return true; }
      if (v.text == "LOCAL" || v.text == "1") { *out = NPairLossParameter::LOCAL; return true; }
      *err = "npair_loss_param: unknown MiningRegion '" + v.text + "'"; return false;
    };
    auto method = [&](NPairLossParameter::MiningMethod* out) {
      static const char* names[5] = {"HARD", "EASY", "RAND", "RELATIVE_HARD", "RELATIVE_EASY"};
      for (int i = 0; i < 5; ++i) if (v.text == names[i] || v.text == std::to_string(i)) { *out = static_cast<NPairLossParameter::MiningMethod>(i); return true; }
      *err = "npair_loss_param: unknown MiningMethod '" + v.text + "'"; return false;
    };
    NPairLossParameter::MiningRegion r; NPairLossParameter::MiningMethod m;
    if (k.text == "margin_ident") { if (!parse_float(v, &f)) { *err = "bad float for margin_ident"; return false; } p->set_margin_ident(f); }
    else if (k.text == "margin_diff") { if (!parse_float(v, &f)) { *err = "bad float for margin_diff"; return false; } p->set_margin_diff(f); }
    else if (k.text == "identsn") { if (!parse_float(v, &f)) { *err = "bad float for identsn"; return false; } p->set_identsn(f); }
    else if (k.text == "diffsn") { if (!parse_float(v, &f)) { *err = "bad float for diffsn"; return false; } p->set_diffsn(f); }
    else if (k.text == "ap_mining_region") { if (!region(&r)) return false; p->set_ap_mining_region(r); }
    else if (k.text == "an_mining_region") { if (!region(&r)) return false; p->set_an_mining_region(r); }
    else if (k.text == "ap_mining_method") { if (!method(&m)) return false; p->set_ap_mining_method(m); }
    else if (k.text == "an_mining_method") { if (!method(&m)) return false; p->set_an_mining_method(m); }
    else { *err = "npair_loss_param: unknown field '" + k.text + "'"; return false; }
  }
}

// nested message of a layer other than npair_loss_param: its scalar fields are kept as text ("message.field"), deeper levels skipped
bool capture_message(Lexer& lx, const std::string& name, LayerParameter* L, std::string* err) {
  for (;;) {
    Tok k = lx.next();
    if (k.kind == Tok::RBRACE) return true;
    if (k.kind == Tok::END) { *err = "unterminated message " + name; return false; }
    if (k.kind != Tok::IDENT) continue;
    Tok c = lx.next();
    if (c.kind == Tok::LBRACE) { if (!skip_message(lx, err)) return false; continue; }
    if (c.kind != Tok::COLON) { *err = name + ": expected ':' after " + k.text; return false; }
    Tok v = lx.next();
    if (v.kind == Tok::LBRACE) { if (!skip_message(lx, err)) return false; continue; }
    L->set_extra(name + "." + k.text, v.text);
  }
}

bool parse_layer(Lexer& lx, LayerParameter* L, std::string* err) {
  for (;;) {
    Tok k = lx.next();
    if (k.kind == Tok::RBRACE) return true;
    if (k.kind == Tok::END) { *err = "unterminated layer block"; return false; }
    if (k.kind != Tok::IDENT) { *err = "layer: expected field name, got '" + k.text + "'"; return false; }
    Tok c = lx.next();
    if (c.kind == Tok::LBRACE) {
      if (k.text == "npair_loss_param") { if (!parse_npair(lx, L->mutable_npair_loss_param(), err)) return false; }
      else if (!capture_message(lx, k.text, L, err)) return false;
      continue;
    }
    if (c.kind != Tok::COLON) { *err = "layer: expected ':' or '{' after " + k.text; return false; }
    Tok v = lx.next();
    if (v.kind == Tok::LBRACE) {   // "field: { ... }" form
      if (k.text == "npair_loss_param") { if (!parse_npair(lx, L->mutable_npair_loss_param(), err)) return false; }
      else if (!capture_message(lx, k.text, L, err)) return false;
      continue;
    }
    if (k.text == "name") L->set_name(v.text);
    else if (k.text == "type") L->set_type(v.text);
    else if (k.text == "bottom") L->add_bottom(v.text);
    else if (k.text == "top") L->add_top(v.text);
    else if (k.text == "loss_weight") { float f; if (!parse_float(v, &f)) { *err = "bad loss_weight"; return false; } L->add_loss_weight(f); }
    // every other scalar field (phase, ...) is irrelevant to this layer
  }
}
}  // namespace

bool ReadLayersFromText(const std::string& text, std::vector<LayerParameter>* layers, std::string* error) {
  Lexer lx(text);
  std::string err;
  for (;;) {
    Tok k = lx.next();
    if (k.kind == Tok::END) return true;
    if (k.kind == Tok::RBRACE) continue;   // stray closers from elided regions (usage/def.prototxt:110-111)
    if (k.kind != Tok::IDENT) continue;
    Tok c = lx.next();
    if (c.kind == Tok::COLON) {
      Tok v = lx.next();
      if (v.kind == Tok::LBRACE && !skip_message(lx, &err)) { if (error) *error = err; return false; }
      continue;
    }
    if (c.kind != Tok::LBRACE) continue;
    if (k.text == "layer" || k.text == "layers") {
      LayerParameter L;
      if (!parse_layer(lx, &L, &err)) { if (error) *error = err; return false; }
      layers->push_back(L);
    } else if (!skip_message(lx, &err)) { if (error) *error = err; return false; }
  }
}

// top-level "key: value" pairs of a prototxt (the solver file, usage/solver.prototxt); nested messages are skipped
bool ReadScalarsFromText(const std::string& text, std::map<std::string, std::string>* out, std::string* error) {
  Lexer lx(text);
  std::string err;
  for (;;) {
    Tok k = lx.next();
    if (k.kind == Tok::END) return true;
    if (k.kind != Tok::IDENT) continue;
    Tok c = lx.next();
    if (c.kind == Tok::LBRACE) { if (!skip_message(lx, &err)) { if (error) *error = err; return false; } continue; }
    if (c.kind != Tok::COLON) continue;
    Tok v = lx.next();
    if (v.kind == Tok::LBRACE) { if (!skip_message(lx, &err)) { if (error) *error = err; return false; } continue; }
    if (!out->count(k.text)) (*out)[k.text] = v.text;
  }
}

}  // namespace caffe
