// tools/emu/simt.cc -- TEST INFRASTRUCTURE ONLY: the SIMT launcher of include/simt.h (one OS thread per warp, 32 coroutine lanes).
#include <cuda_runtime.h>

#if !defined(__x86_64__)
#error "the lane switch of the SIMT emulator is written for x86-64"
#endif
// void emu_switch(void **save_sp, void *load_sp): push the callee-saved registers, park the stack pointer, adopt the
// other one, pop its registers and return into it (System V AMD64; no signal mask, no floating-point state: both are
// the same for every lane)
asm(R"(
.text
.globl emu_switch
.type emu_switch, @function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch, .-emu_switch
)");

// ThreadSanitizer build (build_emu.py --tsan): every lane is a TSan fiber, so that accesses of different WARPS (OS threads) to
// shared memory that are not ordered by an mbarrier / __syncthreads are reported, while the lanes of one warp -- which only
// interleave at the switches below, each a happens-before edge -- are not.
#if defined(__SANITIZE_THREAD__)
extern "C" {
void *__tsan_get_current_fiber(void);
void *__tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void *fiber);
void __tsan_switch_to_fiber(void *fiber, unsigned flags);
}
#define EMU_TSAN 1
#else
#define EMU_TSAN 0
#endif

namespace emu {
thread_local Warp *cur_warp = nullptr;
#if EMU_TSAN
thread_local void *tsan_main = nullptr;
thread_local void *tsan_lane[32] = {};
#endif
void lane_to_main(Warp &w, int me) {
#if EMU_TSAN
  __tsan_switch_to_fiber(tsan_main, 0);
#endif
  emu_switch(&w.lane_sp[me], w.main_sp);
}

namespace {
constexpr size_t kLaneStack = 512 * 1024;

void lane_finished(Warp &w) {  // a lane that has left the kernel no longer takes part in collectives
  w.done[w.cur] = true;
  w.active--;
  w.progress = true;
  if (w.active > 0 && w.arrived == w.active) {
    w.arrived = 0;
    w.gen++;
  }
  if (w.active > 0 && w.sync_arrived == w.active) not_emulated("a thread leaving the kernel while its warp waits at __syncthreads");
}
void lane_entry() {
  Warp &w = *cur_warp;
  (*w.cta->body)();
  lane_finished(w);
  void *dead = nullptr;
#if EMU_TSAN
  __tsan_switch_to_fiber(tsan_main, 0);
#endif
  emu_switch(&dead, w.main_sp);  // for good: a finished lane is never resumed
  std::abort();
}
void run_warp(Warp *wp, int nlanes) {
  Warp &w = *wp;
  cur_warp = wp;
#if EMU_TSAN
  tsan_main = __tsan_get_current_fiber();
#endif
  w.active = nlanes;
  for (int l = 0; l < 32; l++) {
    w.done[l] = l >= nlanes;
    if (l >= nlanes) continue;
    // (stacks are malloc'ed once per (warp, lane) slot and reused by every launch: zero-filling them would dominate)
    // a fresh lane "returns" into lane_entry: six zeroed callee-saved registers, the entry address, a dummy caller
    uintptr_t top = ((uintptr_t)(w.stack[l] + kLaneStack)) & ~(uintptr_t)15;
    void **sp = (void **)top;
    *--sp = nullptr;                 // where lane_entry would return to (it never does)
    *--sp = (void *)&lane_entry;     // popped by emu_switch's ret: rsp is then 8 mod 16, as after a call
    for (int k = 0; k < 6; k++) *--sp = nullptr;
    w.lane_sp[l] = (void *)sp;
#if EMU_TSAN
    tsan_lane[l] = __tsan_create_fiber(0);
#endif
  }
  while (w.active > 0) {
    w.progress = false;
    for (int l = 0; l < 32; l++) {
      if (w.done[l]) continue;
      w.cur = l;
      set_lane_ids(w);
#if EMU_TSAN
      __tsan_switch_to_fiber(tsan_lane[l], 0);
#endif
      emu_switch(&w.main_sp, w.lane_sp[l]);
    }
    if (!w.progress && w.active > 0) std::this_thread::yield();  // every lane waits for another warp: let that one run
  }
#if EMU_TSAN
  for (int l = 0; l < nlanes; l++) __tsan_destroy_fiber(tsan_lane[l]);
#endif
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
        static std::vector<char *> pool;  // lane stacks, by warp * 32 + lane (launches do not overlap)
        while (pool.size() < (size_t)nwarps * 32) pool.push_back(static_cast<char *>(std::malloc(kLaneStack)));
        std::vector<Warp> warps((size_t)nwarps);
        std::vector<std::thread> th;
        for (int k = 0; k < nwarps; k++) {
          warps[(size_t)k].cta = &cta;
          warps[(size_t)k].id = k;
          for (int l = 0; l < 32; l++) warps[(size_t)k].stack[l] = pool[(size_t)k * 32 + l];
          const int nl = (int)std::min(32u, nthreads - (unsigned)k * 32u);
          th.emplace_back(run_warp, &warps[(size_t)k], nl);
        }
        for (auto &t : th) t.join();
      }
}
}  // namespace emu
