// Error state, named-weight lookup and device pools.
#include "model.cuh"

#include <string.h>
#include <stdlib.h>

namespace echo {

static thread_local char g_err[1024] = "";
thread_local int64_t g_launches = 0;

void set_last_error(const char* msg) {
  strncpy(g_err, msg, sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}
const char* last_error() { return g_err; }

void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

bool dbg_skip(const char* name) {
  static const char* env = getenv("ECHO_SKIP");
  if (!env) return false;
  const size_t n = strlen(name);
  for (const char* p = env; (p = strstr(p, name)) != nullptr; p += n)
    if ((p == env || p[-1] == ',') && (p[n] == 0 || p[n] == ',')) return true;
  return false;
}
bool pdl_enabled() {
  static const bool on = getenv("ECHO_NO_PDL") == nullptr;
  return on;
}
bool dbg_trace() {
  static const bool on = getenv("ECHO_TRACE") != nullptr;
  return on;
}

void WeightMap::load(const echo_weight_t* w, int n) {
  ECHO_CHECK(w || n == 0, "weights: null table");
  for (int i = 0; i < n; ++i) {
    ECHO_CHECK(w[i].name && w[i].ndim >= 0 && w[i].ndim <= 6, "weights: bad entry %d", i);
    if (w[i].dtype != 0) continue;   // int64 BatchNorm counters are not used in eval mode
    ECHO_CHECK(w[i].data != nullptr, "weights: null data for %s", w[i].name);
    WView v;
    v.p = (const float*)w[i].data;
    v.shape.assign(w[i].shape, w[i].shape + w[i].ndim);
    m[w[i].name] = v;
  }
}

const WView& WeightMap::get(const std::string& k) const {
  auto it = m.find(k);
  if (it == m.end()) fail(ECHO_ERR_INVALID, "missing weight '%s'", k.c_str());
  return it->second;
}

const WView& WeightMap::get(const std::string& k, std::initializer_list<int64_t> shape) const {
  const WView& v = get(k);
  bool ok = v.shape.size() == shape.size();
  if (ok) {
    size_t i = 0;
    for (auto s : shape) ok = ok && (v.shape[i++] == s);
  }
  if (!ok) {
    std::string got, want;
    for (auto s : v.shape) got += std::to_string(s) + ",";
    for (auto s : shape) want += std::to_string(s) + ",";
    fail(ECHO_ERR_INVALID, "weight '%s' has shape (%s) expected (%s)", k.c_str(), got.c_str(), want.c_str());
  }
  return v;
}

void* DevPool::alloc(size_t bytes) {
  bytes = (bytes + 255) & ~size_t(255);
  if (bytes > left) {
    size_t slab = bytes > (size_t(64) << 20) ? bytes : (size_t(64) << 20);
    void* p = nullptr;
    ECHO_CUDA(cudaMalloc(&p, slab));
    slabs.push_back(p);
    cur = (char*)p;
    left = slab;
    total += slab;
  }
  void* r = cur;
  cur += bytes;
  left -= bytes;
  return r;
}

float* DevPool::upload(const std::vector<float>& h) {
  float* d = alloc_n<float>(h.size());
  ECHO_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}

void DevPool::destroy() {
  for (void* p : slabs) cudaFree(p);
  slabs.clear();
  cur = nullptr;
  left = 0;
}

}  // namespace echo
