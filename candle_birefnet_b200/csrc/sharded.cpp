// Single-process multi-GPU driver behind the C ABI: one model handle + one host thread per GPU,
// weights replicated, a batch split into contiguous image ranges, no collective -- every image's forward is
// independent (eval-mode BatchNorm everywhere: src/decoder.rs:129,139; src/aspp.rs:220,316,330).  Each worker thread
// calls the ordinary brn_forward path of its handle with HOST pointers into the caller's buffers: the H2D copy of a
// shard, its kernels and the D2H copy of its masks run on that GPU's own stream, so the copies of one GPU overlap the
// kernels of the others and nothing crosses between devices.  Results are bit-identical to a single-handle call
// (per-image results do not depend on the batch an image is in -- tests/test_gpu_shard.py).
#include <cstring>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include "model.h"

using namespace brn;

struct brn_sharded {
  std::vector<Model*> models;
  std::vector<int> devices;
  ~brn_sharded();
};
brn_sharded::~brn_sharded() { for (Model* m : models) delete m; }

namespace brn {
brn_sharded* sharded_create(const brn_config& cfg, const int* devices, int n) {
  BRN_CHECK(devices && n > 0 && n <= 64, 1, "brn_sharded_create: need 1..64 devices");
  brn_sharded* s = new brn_sharded;
  try {
    for (int i = 0; i < n; ++i) {
      s->models.push_back(new Model(cfg, devices[i]));
      s->devices.push_back(devices[i]);
    }
  } catch (...) {
    delete s;
    throw;
  }
  return s;
}

// runs fn(shard index) on one thread per handle and rethrows the first failure
template <class F>
static void for_each_handle(brn_sharded* s, F&& fn) {
  const size_t n = s->models.size();
  std::vector<std::exception_ptr> errs(n);
  std::vector<std::thread> th;
  th.reserve(n);
  for (size_t i = 0; i < n; ++i)
    th.emplace_back([&, i] {
      try { fn(i); } catch (...) { errs[i] = std::current_exception(); }
    });
  for (auto& t : th) t.join();
  for (auto& e : errs) if (e) std::rethrow_exception(e);
}

void sharded_set_tensor(brn_sharded* s, const char* key, const void* data, int dtype, const int64_t* shape, int rank) {
  for (Model* m : s->models) {
    std::lock_guard<std::mutex> lk(m->mu);
    m->set_tensor(key, data, dtype, shape, rank);
  }
}

int sharded_load_safetensors(brn_sharded* s, const char* path) {
  int n = 0;
  for (Model* m : s->models) {
    std::lock_guard<std::mutex> lk(m->mu);
    n = load_safetensors(*m, path);
  }
  return n;
}

void sharded_finalize(brn_sharded* s) {
  for_each_handle(s, [&](size_t i) {
    std::lock_guard<std::mutex> lk(s->models[i]->mu);
    s->models[i]->finalize();
  });
}

void sharded_forward(brn_sharded* s, const float* x, int B, int H, int W, float* out, bool apply_sigmoid) {
  BRN_CHECK(x && out && B > 0, 1, "brn_sharded_forward: bad argument");
  const int n = (int)s->models.size();
  // contiguous split; the first B % n shards get one extra image (same rule as candle_birefnet_b200/shard.py)
  std::vector<int> lo(n + 1, 0);
  for (int i = 0; i < n; ++i) lo[i + 1] = lo[i] + B / n + (i < B % n ? 1 : 0);
  for_each_handle(s, [&](size_t i) {
    const int nb = lo[i + 1] - lo[i];
    if (nb <= 0) return;
    s->models[i]->forward(x + (size_t)lo[i] * 3 * H * W, nb, H, W, false, out + (size_t)lo[i] * H * W, false, nullptr,
                          apply_sigmoid);
  });
}
}  // namespace brn
