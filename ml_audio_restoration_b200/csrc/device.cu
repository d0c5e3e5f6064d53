// Per-device host state shared by the launchers: SM count, one-time kernel attributes, the conv shared-memory budget.
// Function attributes (cudaFuncSetAttribute) and the SM count belong to a device / context, not to the process: one
// process may drive several GPUs (RestorationPipeline(device="cuda:1") next to "cuda:0"), so every "done once" flag here
// is kept per device.
#include <atomic>

#include "ar_common.cuh"

namespace ar {

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 ? 0 : dev;
}

bool DeviceOnce::pending() const {
  const int dev = current_device();
  return dev >= MAX_DEVICES || !(bits.load(std::memory_order_acquire) >> dev & 1ull);
}
void DeviceOnce::done() {
  const int dev = current_device();
  if (dev < MAX_DEVICES) bits.fetch_or(1ull << dev, std::memory_order_release);
}

int sm_count() {
  static std::atomic<int> n[DeviceOnce::MAX_DEVICES];   // zero-initialised
  const int dev = current_device();
  int v = dev < DeviceOnce::MAX_DEVICES ? n[dev].load(std::memory_order_relaxed) : 0;
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    if (dev < DeviceOnce::MAX_DEVICES) n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

static std::atomic<int> g_conv_smem{227 * 1024};
int set_conv_smem_kb(int kb) {
  if (kb < 64 || kb > 227) { set_error("conv shared-memory budget must be within [64, 227] KB"); return AR_ERR_INVALID; }
  g_conv_smem.store(kb * 1024);
  return AR_OK;
}
int conv_smem_budget() { return g_conv_smem.load(); }

}  // namespace ar
