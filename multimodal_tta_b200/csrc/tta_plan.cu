// tta_plan: the launch list of one adaptation step as a C object (plan_create / tta_step behind the
// C ABI).  The host side builds the list ONCE per input shape by calling the ordinary tta_* entry points between
// tta_plan_begin(plan, section) and tta_plan_end(): while recording, those calls store themselves instead of
// launching.  Afterwards one call -- tta_plan_run(plan, section, stream), or tta_step(plan, stream) for
// forward + entropy head + backward -- enqueues every kernel of the section on `stream` (capturable into a CUDA
// graph like any other launch sequence).  A host in any language that can call C can therefore drive the step:
// the reference's Python trainer loop (src/core/trainers/seg_trainer.py:97-145) does it through ctypes.
#include <vector>

#include "tta_common.cuh"

struct tta_plan {
  std::vector<std::function<int(cudaStream_t)>> sec[TTA_PLAN_SECTIONS];
};

static thread_local tta_plan* g_rec_plan = nullptr;
static thread_local int g_rec_sec = 0;

bool tta_recording() { return g_rec_plan != nullptr; }

void tta_record_push(std::function<int(cudaStream_t)> fn) { g_rec_plan->sec[g_rec_sec].push_back(std::move(fn)); }

extern "C" {

int tta_plan_create(tta_plan** out) {
  TTA_REQUIRE(out != nullptr, "tta_plan_create: null pointer");
  *out = new tta_plan();
  return TTA_OK;
}

int tta_plan_destroy(tta_plan* plan) {
  if (g_rec_plan == plan) g_rec_plan = nullptr;
  delete plan;
  return TTA_OK;
}

// start recording into `section` (0 forward, 1 training head, 2 inference head, 3 backward, 4..7 free); the section's
// previous content is dropped
int tta_plan_begin(tta_plan* plan, int section) {
  TTA_REQUIRE(plan != nullptr && section >= 0 && section < TTA_PLAN_SECTIONS, "tta_plan_begin: bad plan / section %d", section);
  TTA_REQUIRE(g_rec_plan == nullptr, "tta_plan_begin: a recording is already active on this thread");
  plan->sec[section].clear();
  g_rec_plan = plan;
  g_rec_sec = section;
  return TTA_OK;
}

int tta_plan_end(void) {
  TTA_REQUIRE(g_rec_plan != nullptr, "tta_plan_end: no recording is active on this thread");
  g_rec_plan = nullptr;
  return TTA_OK;
}

int tta_plan_num_launches(const tta_plan* plan, int section) {
  if (plan == nullptr || section < 0 || section >= TTA_PLAN_SECTIONS) return -1;
  return (int)plan->sec[section].size();
}

int tta_plan_run(const tta_plan* plan, int section, cudaStream_t stream) {
  TTA_REQUIRE(plan != nullptr && section >= 0 && section < TTA_PLAN_SECTIONS, "tta_plan_run: bad plan / section %d", section);
  TTA_REQUIRE(g_rec_plan == nullptr, "tta_plan_run: a recording is active on this thread");
  for (const auto& fn : plan->sec[section]) {
    const int rc = fn(stream);
    if (rc != TTA_OK) return rc;   // the failing call has set the thread's error message
  }
  return TTA_OK;
}

// forward + training head (loss, dlogits) + backward of the recorded network on whatever the input buffers hold
int tta_step(const tta_plan* plan, cudaStream_t stream) {
  int rc = tta_plan_run(plan, 0, stream);
  if (rc == TTA_OK) rc = tta_plan_run(plan, 1, stream);
  if (rc == TTA_OK) rc = tta_plan_run(plan, 3, stream);
  return rc;
}

// bytes of the scratch workspace the norm / head kernels of a layer shape need (what the *_workspace_floats queries
// return, in bytes): [1024 counters][per-block partial sums]
long long tta_workspace_bytes(int N, int C8, long long V) { return 4 * tta_norm_workspace_floats(N, C8, V); }

}  // extern "C"
