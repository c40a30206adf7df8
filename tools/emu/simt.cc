// tools/emu/simt.cc -- TEST INFRASTRUCTURE ONLY: the SIMT launcher of include/simt.h (one OS thread per warp, 32 coroutine lanes).
#include <cuda_runtime.h>

namespace emu {
thread_local Warp *cur_warp = nullptr;

namespace {
constexpr size_t kLaneStack = 512 * 1024;

void lane_finished(Warp &w) {  // a lane that has left the kernel no longer takes part in collectives
  w.done[w.cur] = true;
  w.active--;
  if (w.active > 0 && w.arrived == w.active) {
    w.arrived = 0;
    w.gen++;
  }
  if (w.active > 0 && w.sync_arrived == w.active) not_emulated("a thread leaving the kernel while its warp waits at __syncthreads");
}
void trampoline() {
  Warp &w = *cur_warp;
  (*w.cta->body)();
  lane_finished(w);
}
void run_warp(Warp *wp, int nlanes) {
  Warp &w = *wp;
  cur_warp = wp;
  w.active = nlanes;
  for (int l = 0; l < 32; l++) {
    w.done[l] = l >= nlanes;
    if (l >= nlanes) continue;
    w.stack[l].resize(kLaneStack);
    getcontext(&w.lane[l]);
    w.lane[l].uc_stack.ss_sp = w.stack[l].data();
    w.lane[l].uc_stack.ss_size = kLaneStack;
    w.lane[l].uc_link = &w.main;
    makecontext(&w.lane[l], trampoline, 0);
  }
  while (w.active > 0)
    for (int l = 0; l < 32; l++) {
      if (w.done[l]) continue;
      w.cur = l;
      set_lane_ids(w);
      swapcontext(&w.main, &w.lane[l]);
    }
  cur_warp = nullptr;
}
}  // namespace

void launch_simt(dim3 grid, dim3 block, size_t shmem, std::function<void()> body) {
  const unsigned nthreads = block.x * block.y * block.z;
  const int nwarps = (int)((nthreads + 31) / 32);
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        Cta cta;
        cta.block = block;
        cta.grid = grid;
        cta.block_idx = Idx{bx, by, bz};
        cta.smem.assign(shmem + 128, 0xCD);  // uninitialised shared memory is not zero
        cta.body = &body;
        cta.nwarps = nwarps;
        std::vector<Warp> warps((size_t)nwarps);
        std::vector<std::thread> th;
        for (int k = 0; k < nwarps; k++) {
          warps[(size_t)k].cta = &cta;
          warps[(size_t)k].id = k;
          const int nl = (int)std::min(32u, nthreads - (unsigned)k * 32u);
          th.emplace_back(run_warp, &warps[(size_t)k], nl);
        }
        for (auto &t : th) t.join();
      }
}
}  // namespace emu
