// Library-level entry points + shared host-side validation.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace advs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static_assert(sizeof(advs_conv_params) == 208, "advs_conv_params layout changed: update _capi.ConvParams and the tests");
constexpr int kMaxDevices = 64;
static std::atomic<unsigned char> g_once[kOnceSlots][kMaxDevices];
static std::atomic<int> g_sms[kMaxDevices];

static int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev % kMaxDevices;
}

bool first_use_on_device(int slot) { return g_once[slot][current_device_slot()].exchange(1) == 0; }
void forget_first_use(int slot) { g_once[slot][current_device_slot()].store(0); }

int current_device_sms() {
  const int d = current_device_slot();
  int n = g_sms[d].load();
  if (n <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sms[d].store(n);
  }
  return n;
}

// -1: follow the environment (ADVS_PDL=1 turns it on; default off); 0 / 1: set by advs_set_pdl
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  static const bool env_on = [] { const char* e = getenv("ADVS_PDL"); return e && atoi(e) != 0; }();
  const int v = g_pdl.load(std::memory_order_relaxed);
  return v < 0 ? env_on : v != 0;
}

int validate_conv(const advs_conv_params* p, const char* who) {
  ADVS_CHECK_ARG(p != nullptr, "%s: null params", who);
  ADVS_CHECK_ARG(p->B > 0 && p->H > 0 && p->W > 0 && p->Cout > 0, "%s: bad output shape", who);
  ADVS_CHECK_ARG(p->stride == 1 || p->stride == 2, "%s: stride must be 1 or 2", who);
  ADVS_CHECK_ARG(p->nseg >= 1 && p->nseg <= 3, "%s: nseg must be 1..3", who);
  for (int s = 0; s < p->nseg; ++s) {
    ADVS_CHECK_ARG(p->seg[s].x && p->seg[s].w, "%s: segment %d has null pointers", who, s);
    ADVS_CHECK_ARG(p->seg[s].taps == 9 || p->seg[s].taps == 1 || (s == 0 && p->seg[s].taps == 4),
                   "%s: taps must be 9, 1 or (segment 0 of an upsample phase) 4", who);
    ADVS_CHECK_ARG(p->seg[s].C > 0 && p->seg[s].C % 4 == 0, "%s: C must be a multiple of 4", who);
    ADVS_CHECK_ARG(s == 0 || p->seg[s].taps == 1, "%s: shortcut segments must be 1x1", who);
  }
  ADVS_CHECK_ARG(p->stride == 1 || p->seg[0].taps == 9, "%s: stride 2 needs a 3x3 kernel", who);
  ADVS_CHECK_ARG(p->up_phase >= 0 && p->up_phase <= 4, "%s: up_phase must be 0..4", who);
  ADVS_CHECK_ARG((p->up_phase != 0) == (p->seg[0].taps == 4), "%s: taps = 4 goes with up_phase != 0", who);
  ADVS_CHECK_ARG(p->up_phase == 0 || (p->stride == 1 && p->out_mode == 0 && p->nseg == 1 && !p->residual),
                 "%s: an upsample phase is a plain stride-1 NHWC convolution", who);
  ADVS_CHECK_ARG(p->dtype == ADVS_F32 || p->dtype == ADVS_BF16, "%s: bad dtype", who);
  ADVS_CHECK_ARG(!p->y_lo || (p->out_mode == 0 && p->dtype == ADVS_BF16 && ((uintptr_t)p->y_lo % 32) == 0 && p->Cout % 32 == 0),
                 "%s: y_lo needs out_mode 0, bf16, Cout %% 32 == 0 and a 32-byte aligned pointer", who);
  ADVS_CHECK_ARG(p->operand_f16 == 0 || p->dtype == ADVS_BF16, "%s: operand_f16 goes with the bf16 (16-bit) mode", who);
  if (p->out_mode == 0) {
    ADVS_CHECK_ARG(p->y != nullptr, "%s: y is null", who);
  } else if (p->out_mode == 1) {
    ADVS_CHECK_ARG(p->q && p->k && p->vt, "%s: q/k/vt null in qkv mode", who);
    ADVS_CHECK_ARG(p->heads > 0 && p->Cout % (3 * p->heads) == 0, "%s: Cout not 3*heads*dh", who);
    ADVS_CHECK_ARG(p->residual == nullptr, "%s: residual unsupported in qkv mode", who);
  } else if (p->out_mode == 2) {
    ADVS_CHECK_ARG(p->y != nullptr, "%s: y is null", who);
    ADVS_CHECK_ARG(p->cout_valid > 0 && p->cout_valid <= p->Cout, "%s: cout_valid out of range", who);
    ADVS_CHECK_ARG(p->residual == nullptr, "%s: residual unsupported in NCHW output mode", who);
  } else {
    ADVS_CHECK_ARG(false, "%s: bad out_mode", who);
  }
  return ADVS_OK;
}

}  // namespace advs

extern "C" {

int advs_version(void) { return 100; }

int advs_set_pdl(int on) { return advs::g_pdl.exchange(on < 0 ? -1 : (on ? 1 : 0)); }

const char* advs_last_error(void) { return advs::g_err; }

int advs_device_is_sm100(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

}  // extern "C"
