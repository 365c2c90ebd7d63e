// weights.cc — safetensors reader + BatchNorm folding for the policy/value net.
//
// Replaces the weight hand-off of the reference: VarStore::save (ref: src/learner.rs:192,
// src/learner_concurrent.rs:155-156) -> VarStore::load (ref: src/main.rs:61).  The network
// architecture is model/mod.rs:152-184 + model/connect_four.rs:50-73 (tictactoe.rs:50-73).
//
// Two naming schemes are accepted:
//  (1) explicit: conv{i}.{weight,bias}, bn{i}.{weight,bias,running_mean,running_var} for
//      i = 0 (stem), 1..8 (residual convs), 9 (policy conv), 10 (value conv);
//      policy_fc.{weight,bias}, value_fc.{weight,bias}.
//  (2) tch VarStore root-path names: every layer is created on the SAME nn::Path (model/mod.rs:155-159,
//      connect_four.rs:55-71), so tch de-duplicates names by appending "__{n}" where n is the number
//      of variables created so far.  Tensors are therefore matched by CREATION ORDER (the numeric
//      suffix) within each (name, rank) class.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "chess_net.cuh"
#include "evaluator.cuh"

namespace spb {
namespace {

struct TensorInfo {
  std::string name;
  std::string dtype;
  std::vector<int64_t> shape;
  size_t begin = 0, end = 0;
};

// Minimal JSON reader for the safetensors header: {"name": {"dtype": "..", "shape": [..], "data_offsets": [b, e]}, ...}
struct JsonCursor {
  const char* p;
  const char* e;
  bool fail = false;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool eat(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
  std::string str() {
    ws();
    std::string s;
    if (p >= e || *p != '"') { fail = true; return s; }
    ++p;
    while (p < e && *p != '"') {
      if (*p == '\\' && p + 1 < e) { ++p; }
      s.push_back(*p++);
    }
    if (p >= e) { fail = true; return s; }
    ++p;
    return s;
  }
  int64_t num() {
    ws();
    int64_t v = 0;
    bool any = false, neg = false;
    if (p < e && *p == '-') { neg = true; ++p; }
    while (p < e && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); ++p; any = true; }
    if (!any) fail = true;
    return neg ? -v : v;
  }
  void skip_value() {   // skips any JSON value (used for __metadata__)
    ws();
    if (p >= e) { fail = true; return; }
    if (*p == '"') { str(); return; }
    if (*p == '{' || *p == '[') {
      char open = *p, close = (open == '{') ? '}' : ']';
      int depth = 0;
      while (p < e) {
        if (*p == '"') { str(); continue; }
        if (*p == open) ++depth;
        else if (*p == close) { --depth; if (depth == 0) { ++p; return; } }
        ++p;
      }
      fail = true;
      return;
    }
    while (p < e && *p != ',' && *p != '}' && *p != ']') ++p;
  }
};

bool parse_header(const uint8_t* blob, size_t n, std::vector<TensorInfo>* out, size_t* data_off, std::string* err) {
  if (n < 8) { *err = "blob shorter than the 8-byte header length"; return false; }
  uint64_t hlen = 0;
  std::memcpy(&hlen, blob, 8);
  if (hlen > n - 8 || hlen > (1ull << 26)) { *err = "header length out of range"; return false; }
  JsonCursor c{reinterpret_cast<const char*>(blob + 8), reinterpret_cast<const char*>(blob + 8 + hlen)};
  *data_off = 8 + (size_t)hlen;
  if (!c.eat('{')) { *err = "header is not a JSON object"; return false; }
  if (c.eat('}')) return true;
  do {
    std::string name = c.str();
    if (c.fail || !c.eat(':')) { *err = "malformed header near key"; return false; }
    if (name == "__metadata__") { c.skip_value(); if (c.fail) { *err = "malformed __metadata__"; return false; } continue; }
    TensorInfo t;
    t.name = name;
    if (!c.eat('{')) { *err = "tensor entry is not an object: " + name; return false; }
    do {
      std::string k = c.str();
      if (c.fail || !c.eat(':')) { *err = "malformed tensor entry: " + name; return false; }
      if (k == "dtype") t.dtype = c.str();
      else if (k == "shape") {
        if (!c.eat('[')) { *err = "shape is not an array: " + name; return false; }
        if (!c.eat(']')) { do { t.shape.push_back(c.num()); } while (c.eat(',')); if (!c.eat(']')) { *err = "malformed shape: " + name; return false; } }
      } else if (k == "data_offsets") {
        if (!c.eat('[')) { *err = "data_offsets is not an array: " + name; return false; }
        t.begin = (size_t)c.num();
        if (!c.eat(',')) { *err = "malformed data_offsets: " + name; return false; }
        t.end = (size_t)c.num();
        if (!c.eat(']')) { *err = "malformed data_offsets: " + name; return false; }
      } else c.skip_value();
      if (c.fail) { *err = "malformed tensor entry: " + name; return false; }
    } while (c.eat(','));
    if (!c.eat('}')) { *err = "unterminated tensor entry: " + name; return false; }
    out->push_back(t);
  } while (c.eat(','));
  if (!c.eat('}')) { *err = "unterminated header"; return false; }
  return true;
}

struct Loaded {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

bool load_f32(const uint8_t* blob, size_t n, size_t data_off, const TensorInfo& t, Loaded* out, std::string* err) {
  size_t count = 1;
  for (int64_t d : t.shape) { if (d < 0) { *err = "negative dim: " + t.name; return false; } count *= (size_t)d; }
  size_t esz = t.dtype == "F32" ? 4 : (t.dtype == "F64" ? 8 : (t.dtype == "BF16" || t.dtype == "F16" ? 2 : 0));
  if (esz == 0 || t.dtype == "F16") { *err = "unsupported dtype " + t.dtype + " for " + t.name + " (expected F32)"; return false; }
  if (t.end < t.begin || t.end - t.begin != count * esz || data_off + t.end > n) { *err = "data_offsets do not match shape: " + t.name; return false; }
  const uint8_t* src = blob + data_off + t.begin;
  out->shape = t.shape;
  out->data.resize(count);
  if (t.dtype == "F32") std::memcpy(out->data.data(), src, count * 4);
  else if (t.dtype == "F64") { for (size_t i = 0; i < count; ++i) { double d; std::memcpy(&d, src + i * 8, 8); out->data[i] = (float)d; } }
  else { for (size_t i = 0; i < count; ++i) { uint16_t h; std::memcpy(&h, src + i * 2, 2); uint32_t u = (uint32_t)h << 16; std::memcpy(&out->data[i], &u, 4); } }
  return true;
}

bool shape_is(const Loaded& l, std::initializer_list<int64_t> s) { return l.shape == std::vector<int64_t>(s); }

// splits "weight__12" into ("weight", 12); no suffix -> index -1 (the first variable of that name).
void split_suffix(const std::string& name, std::string* base, long* idx) {
  size_t pos = name.rfind("__");
  *base = name;
  *idx = -1;
  if (pos != std::string::npos && pos + 2 < name.size()) {
    bool digits = true;
    for (size_t i = pos + 2; i < name.size(); ++i) digits &= (name[i] >= '0' && name[i] <= '9');
    if (digits) { *base = name.substr(0, pos); *idx = std::stol(name.substr(pos + 2)); }
  }
  // tolerate a leading path component ("net.weight__3")
  size_t dot = base->rfind('.');
  if (dot != std::string::npos) *base = base->substr(dot + 1);
}

}  // namespace

bool parse_safetensors_net(const void* blob_v, size_t n, int game, HostNet* out, std::string* err) {
  const uint8_t* blob = static_cast<const uint8_t*>(blob_v);
  std::vector<TensorInfo> infos;
  size_t data_off = 0;
  if (!parse_header(blob, n, &infos, &data_off, err)) return false;
  const int R = game == SPB_GAME_CONNECT4 ? 6 : 3, C = game == SPB_GAME_CONNECT4 ? 7 : 3, A = game == SPB_GAME_CONNECT4 ? 7 : 9;
  const int P = R * C;
  out->game = game; out->rows = R; out->cols = C; out->actions = A;

  std::map<std::string, const TensorInfo*> by_name;
  for (const auto& t : infos) by_name[t.name] = &t;

  struct Raw { Loaded w, b, g, beta, mean, var; } conv[NET_CONVS];
  Loaded pfc_w, pfc_b, vfc_w, vfc_b;

  auto get = [&](const std::string& name, Loaded* dst) -> bool {
    auto it = by_name.find(name);
    if (it == by_name.end()) { *err = "missing tensor " + name; return false; }
    return load_f32(blob, n, data_off, *it->second, dst, err);
  };

  if (by_name.count("conv0.weight")) {
    for (int i = 0; i < NET_CONVS; ++i) {
      std::string c = "conv" + std::to_string(i), b = "bn" + std::to_string(i);
      if (!get(c + ".weight", &conv[i].w) || !get(c + ".bias", &conv[i].b) || !get(b + ".weight", &conv[i].g) ||
          !get(b + ".bias", &conv[i].beta) || !get(b + ".running_mean", &conv[i].mean) || !get(b + ".running_var", &conv[i].var))
        return false;
    }
    if (!get("policy_fc.weight", &pfc_w) || !get("policy_fc.bias", &pfc_b) || !get("value_fc.weight", &vfc_w) || !get("value_fc.bias", &vfc_b))
      return false;
  } else {
    // tch creation-order scheme
    struct Item { long idx; const TensorInfo* t; };
    std::vector<Item> w4, w1, w2, bias, rmean, rvar;
    for (const auto& t : infos) {
      std::string base; long idx;
      split_suffix(t.name, &base, &idx);
      if (base == "weight") {
        if (t.shape.size() == 4) w4.push_back({idx, &t});
        else if (t.shape.size() == 1) w1.push_back({idx, &t});
        else if (t.shape.size() == 2) w2.push_back({idx, &t});
      } else if (base == "bias") bias.push_back({idx, &t});
      else if (base == "running_mean") rmean.push_back({idx, &t});
      else if (base == "running_var") rvar.push_back({idx, &t});
    }
    auto by_idx = [](const Item& a, const Item& b) { return a.idx < b.idx; };
    for (auto* v : {&w4, &w1, &w2, &bias, &rmean, &rvar}) std::stable_sort(v->begin(), v->end(), by_idx);
    if (w4.size() != NET_CONVS || w1.size() != NET_CONVS || rmean.size() != NET_CONVS || rvar.size() != NET_CONVS ||
        w2.size() != 2 || bias.size() != 2 * NET_CONVS + 2) {
      *err = "unexpected tensor census for the 4x64 ResNet: conv weights " + std::to_string(w4.size()) + ", bn weights " +
             std::to_string(w1.size()) + ", linear weights " + std::to_string(w2.size()) + ", biases " + std::to_string(bias.size()) +
             ", running_mean " + std::to_string(rmean.size()) + ", running_var " + std::to_string(rvar.size());
      return false;
    }
    // bias order: conv0, bn0, ..., conv8, bn8, conv9, bn9, policy_fc, conv10, bn10, value_fc
    int bi = 0;
    for (int i = 0; i < NET_CONVS; ++i) {
      if (!load_f32(blob, n, data_off, *w4[i].t, &conv[i].w, err) || !load_f32(blob, n, data_off, *w1[i].t, &conv[i].g, err) ||
          !load_f32(blob, n, data_off, *rmean[i].t, &conv[i].mean, err) || !load_f32(blob, n, data_off, *rvar[i].t, &conv[i].var, err))
        return false;
      if (!load_f32(blob, n, data_off, *bias[bi++].t, &conv[i].b, err) || !load_f32(blob, n, data_off, *bias[bi++].t, &conv[i].beta, err))
        return false;
      if (i == 9) { if (!load_f32(blob, n, data_off, *bias[bi++].t, &pfc_b, err)) return false; }
      if (i == 10) { if (!load_f32(blob, n, data_off, *bias[bi++].t, &vfc_b, err)) return false; }
    }
    if (!load_f32(blob, n, data_off, *w2[0].t, &pfc_w, err) || !load_f32(blob, n, data_off, *w2[1].t, &vfc_w, err)) return false;
  }

  // shape checks + BN folding (eval mode): y = (conv(x)+b - mean) * g / sqrt(var + eps) + beta
  for (int i = 0; i < NET_CONVS; ++i) {
    const int ic = i == 0 ? 3 : NET_HIDDEN;
    const int oc = i <= 8 ? NET_HIDDEN : (i == 9 ? NET_POLICY_CH : NET_VALUE_CH);
    if (!shape_is(conv[i].w, {oc, ic, 3, 3}) || !shape_is(conv[i].b, {oc}) || !shape_is(conv[i].g, {oc}) || !shape_is(conv[i].beta, {oc}) ||
        !shape_is(conv[i].mean, {oc}) || !shape_is(conv[i].var, {oc})) {
      *err = "shape mismatch in conv/bn layer " + std::to_string(i);
      return false;
    }
    HostNet::Conv& h = out->conv[i];
    h.oc = oc; h.ic = ic;
    h.w.resize((size_t)oc * ic * 9);
    h.b.resize(oc);
    for (int o = 0; o < oc; ++o) {
      double s = (double)conv[i].g.data[o] / std::sqrt((double)conv[i].var.data[o] + (double)BN_EPS);
      for (int k = 0; k < ic * 9; ++k) h.w[(size_t)o * ic * 9 + k] = (float)((double)conv[i].w.data[(size_t)o * ic * 9 + k] * s);
      h.b[o] = (float)(((double)conv[i].b.data[o] - (double)conv[i].mean.data[o]) * s + (double)conv[i].beta.data[o]);
    }
  }
  if (!shape_is(pfc_w, {A, NET_POLICY_CH * P}) || !shape_is(pfc_b, {A}) || !shape_is(vfc_w, {1, NET_VALUE_CH * P}) || !shape_is(vfc_b, {1})) {
    *err = "shape mismatch in the policy/value linear layers";
    return false;
  }
  out->pfc_w = pfc_w.data; out->pfc_b = pfc_b.data; out->vfc_w = vfc_w.data; out->vfc_b = vfc_b.data;
  return true;
}


// ---- chess network (ref: src/model/chess.rs:50-73) -------------------------------------------------------------------
// Explicit names: conv{i}.{weight,bias} + bn{i}.{weight,bias,running_mean,running_var} for i = 0 (stem), 1..20 (residual
// convs); policy_conv1 / policy_conv2 / value_conv / value_fc1 / value_fc2 .{weight,bias}.  tch VarStore names: creation
// order as for the small nets — torso (conv, bn) x 21, policy convs, value conv, the two Linear layers.
namespace chess {

bool parse_safetensors_chess(const void* blob_v, size_t n, HostNet* out, std::string* err) {
  const uint8_t* blob = static_cast<const uint8_t*>(blob_v);
  std::vector<TensorInfo> infos;
  size_t data_off = 0;
  if (!parse_header(blob, n, &infos, &data_off, err)) return false;
  std::map<std::string, const TensorInfo*> by_name;
  for (const auto& t : infos) by_name[t.name] = &t;
  struct Raw { Loaded w, b, g, beta, mean, var; } conv[NET_CONV3];
  Loaded p1w, p1b, p2w, p2b, vw, vb, f1w, f1b, f2w, f2b;
  auto get = [&](const std::string& name, Loaded* dst) -> bool {
    auto it = by_name.find(name);
    if (it == by_name.end()) { *err = "missing tensor " + name; return false; }
    return load_f32(blob, n, data_off, *it->second, dst, err);
  };
  if (by_name.count("conv0.weight")) {
    for (int i = 0; i < NET_CONV3; ++i) {
      const std::string c = "conv" + std::to_string(i), b = "bn" + std::to_string(i);
      if (!get(c + ".weight", &conv[i].w) || !get(c + ".bias", &conv[i].b) || !get(b + ".weight", &conv[i].g) ||
          !get(b + ".bias", &conv[i].beta) || !get(b + ".running_mean", &conv[i].mean) || !get(b + ".running_var", &conv[i].var))
        return false;
    }
    if (!get("policy_conv1.weight", &p1w) || !get("policy_conv1.bias", &p1b) || !get("policy_conv2.weight", &p2w) ||
        !get("policy_conv2.bias", &p2b) || !get("value_conv.weight", &vw) || !get("value_conv.bias", &vb) ||
        !get("value_fc1.weight", &f1w) || !get("value_fc1.bias", &f1b) || !get("value_fc2.weight", &f2w) || !get("value_fc2.bias", &f2b))
      return false;
  } else {
    struct Item { long idx; const TensorInfo* t; };
    std::vector<Item> w4, w1, w2, bias, rmean, rvar;
    for (const auto& t : infos) {
      std::string base; long idx;
      split_suffix(t.name, &base, &idx);
      if (base == "weight") {
        if (t.shape.size() == 4) w4.push_back({idx, &t});
        else if (t.shape.size() == 1) w1.push_back({idx, &t});
        else if (t.shape.size() == 2) w2.push_back({idx, &t});
      } else if (base == "bias") bias.push_back({idx, &t});
      else if (base == "running_mean") rmean.push_back({idx, &t});
      else if (base == "running_var") rvar.push_back({idx, &t});
    }
    auto by_idx = [](const Item& a, const Item& b) { return a.idx < b.idx; };
    for (auto* v : {&w4, &w1, &w2, &bias, &rmean, &rvar}) std::stable_sort(v->begin(), v->end(), by_idx);
    if (w4.size() != NET_CONV3 + 3 || w1.size() != NET_CONV3 || rmean.size() != NET_CONV3 || rvar.size() != NET_CONV3 || w2.size() != 2 ||
        bias.size() != 2 * NET_CONV3 + 5) {
      *err = "unexpected tensor census for the chess 10x256 ResNet: conv weights " + std::to_string(w4.size()) + ", bn weights " +
             std::to_string(w1.size()) + ", linear weights " + std::to_string(w2.size()) + ", biases " + std::to_string(bias.size());
      return false;
    }
    int bi = 0;
    for (int i = 0; i < NET_CONV3; ++i) {
      if (!load_f32(blob, n, data_off, *w4[i].t, &conv[i].w, err) || !load_f32(blob, n, data_off, *w1[i].t, &conv[i].g, err) ||
          !load_f32(blob, n, data_off, *rmean[i].t, &conv[i].mean, err) || !load_f32(blob, n, data_off, *rvar[i].t, &conv[i].var, err) ||
          !load_f32(blob, n, data_off, *bias[bi++].t, &conv[i].b, err) || !load_f32(blob, n, data_off, *bias[bi++].t, &conv[i].beta, err))
        return false;
    }
    if (!load_f32(blob, n, data_off, *w4[NET_CONV3].t, &p1w, err) || !load_f32(blob, n, data_off, *bias[bi++].t, &p1b, err) ||
        !load_f32(blob, n, data_off, *w4[NET_CONV3 + 1].t, &p2w, err) || !load_f32(blob, n, data_off, *bias[bi++].t, &p2b, err) ||
        !load_f32(blob, n, data_off, *w4[NET_CONV3 + 2].t, &vw, err) || !load_f32(blob, n, data_off, *bias[bi++].t, &vb, err) ||
        !load_f32(blob, n, data_off, *w2[0].t, &f1w, err) || !load_f32(blob, n, data_off, *bias[bi++].t, &f1b, err) ||
        !load_f32(blob, n, data_off, *w2[1].t, &f2w, err) || !load_f32(blob, n, data_off, *bias[bi++].t, &f2b, err))
      return false;
  }
  for (int i = 0; i < NET_CONV3; ++i) {
    const int ic = i == 0 ? NET_IN : NET_HIDDEN, oc = NET_HIDDEN;
    if (!shape_is(conv[i].w, {oc, ic, 3, 3}) || !shape_is(conv[i].b, {oc}) || !shape_is(conv[i].g, {oc}) || !shape_is(conv[i].beta, {oc}) ||
        !shape_is(conv[i].mean, {oc}) || !shape_is(conv[i].var, {oc})) {
      *err = "shape mismatch in conv/bn layer " + std::to_string(i) + " of the chess net";
      return false;
    }
    HostNet::Conv& h = out->conv3[i];
    h.oc = oc; h.ic = ic; h.k = 3;
    h.w.resize((size_t)oc * ic * 9);
    h.b.resize(oc);
    for (int o = 0; o < oc; ++o) {
      const double s = (double)conv[i].g.data[o] / std::sqrt((double)conv[i].var.data[o] + (double)NET_BN_EPS);
      for (int k = 0; k < ic * 9; ++k) h.w[(size_t)o * ic * 9 + k] = (float)((double)conv[i].w.data[(size_t)o * ic * 9 + k] * s);
      h.b[o] = (float)(((double)conv[i].b.data[o] - (double)conv[i].mean.data[o]) * s + (double)conv[i].beta.data[o]);
    }
  }
  if (!shape_is(p1w, {NET_HIDDEN, NET_HIDDEN, 1, 1}) || !shape_is(p1b, {NET_HIDDEN}) || !shape_is(p2w, {NET_MOVE_PLANES, NET_HIDDEN, 1, 1}) ||
      !shape_is(p2b, {NET_MOVE_PLANES}) || !shape_is(vw, {1, NET_HIDDEN, 1, 1}) || !shape_is(vb, {1}) || !shape_is(f1w, {256, 64}) ||
      !shape_is(f1b, {256}) || !shape_is(f2w, {1, 256}) || !shape_is(f2b, {1})) {
    *err = "shape mismatch in the heads of the chess net";
    return false;
  }
  auto set1x1 = [](HostNet::Conv& h, const Loaded& w, const Loaded& b, int oc) { h.oc = oc; h.ic = NET_HIDDEN; h.k = 1; h.w = w.data; h.b = b.data; };
  set1x1(out->p1, p1w, p1b, NET_HIDDEN);
  set1x1(out->p2, p2w, p2b, NET_MOVE_PLANES);
  set1x1(out->vconv, vw, vb, 1);
  out->fc1_w = f1w.data; out->fc1_b = f1b.data; out->fc2_w = f2w.data; out->fc2_b = f2b.data;
  return true;
}

}  // namespace chess
}  // namespace spb
