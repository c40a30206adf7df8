// tools/emu/include/simt.h -- TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory).
//
// A small SIMT emulator for the kernels that cooperate: warp shuffles / votes, __syncwarp, __syncthreads, and the
// sm_90+ machinery the fused step is built on (mbarriers with transaction counts, bulk global->shared copies).
// A CTA runs as one OS thread per warp; the 32 lanes of a warp are coroutines of that thread (a dozen instructions of
// x86-64 stack switching, simt.cc), switched at the collective operations, so a shuffle costs a few user-space switches.  CTAs run one after the other.
// mbarrier waits spin cooperatively (lane yield + sched_yield): a wait that the hardware would satisfy is satisfied here,
// a deadlock shows up as a hang (tests run the emulation under a timeout).  Bulk copies complete at once.
// Nothing here models timing, memory-ordering subtleties of the async proxy, or bank conflicts: it checks LOGIC.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

extern "C" void emu_switch(void **save_sp, void *load_sp);  // saves the callee-saved registers and swaps stacks

namespace emu {

struct Cta;
struct Warp {
  Cta *cta = nullptr;
  int id = 0;
  void *main_sp = nullptr;     // saved stack pointers of the scheduler and of the lanes (emu_switch)
  void *lane_sp[32] = {};
  char *stack[32] = {};
  bool done[32] = {};
  int active = 32, cur = 0;
  // collectives
  uint64_t buf[2][32] = {};
  int arrived = 0;
  unsigned gen = 0;
  int sync_arrived = 0;  // __syncthreads arrivals of this warp's lanes
  unsigned sync_gen = 0;
  bool progress = false;  // some lane got past a wait (or left the kernel) since the scheduler last looked
};
struct MBar {
  unsigned expected = 0;
  int pending = 0;
  long tx = 0;
  std::atomic<unsigned> phase{0};  // read without the lock by the waiters (they poll)
};
struct Cta {
  dim3 block, grid;
  Idx block_idx;
  std::vector<unsigned char> smem;
  std::function<void()> *body = nullptr;
  int nwarps = 0;
  // CTA barrier over warps
  std::mutex mu;
  std::condition_variable cv;
  int bar_count = 0;
  unsigned bar_gen = 0;
  std::map<unsigned, MBar> mbars;  // by shared-memory offset
};

extern thread_local Warp *cur_warp;   // the warp this OS thread runs (null outside a SIMT launch)
constexpr unsigned kSmemBase = 0x1000;  // shared-window addresses start here (0 stays invalid)

inline Warp &warp() {
  if (!cur_warp) not_emulated("a cooperative intrinsic outside a SIMT launch");
  return *cur_warp;
}
void lane_to_main(Warp &w, int me);  // simt.cc: the switch back to the warp's scheduler
inline void set_lane_ids(Warp &w) {
  const unsigned t = (unsigned)(w.id * 32 + w.cur);
  thread_idx = Idx{t % w.cta->block.x, (t / w.cta->block.x) % w.cta->block.y, t / (w.cta->block.x * w.cta->block.y)};
  block_idx = w.cta->block_idx;
  block_dim = w.cta->block;
  grid_dim = w.cta->grid;
}
inline void lane_yield() {  // back to the warp's scheduler
  Warp &w = warp();
  const int me = w.cur;
  lane_to_main(w, me);
  w.cur = me;
  set_lane_ids(w);
}
inline unsigned char *dyn_smem() { return warp().cta->smem.data(); }
inline unsigned smem_u32(const void *p) { return kSmemBase + (unsigned)((const unsigned char *)p - warp().cta->smem.data()); }
inline unsigned char *smem_ptr(unsigned a) { return warp().cta->smem.data() + (a - kSmemBase); }

// ---- warp collectives: every active lane calls; the last one to arrive opens the generation ----
inline void warp_collect(Warp &w, uint64_t mine) {
  const unsigned g = w.gen;
  w.buf[g & 1][w.cur] = mine;
  if (++w.arrived == w.active) {
    w.arrived = 0;
    w.gen++;
  }
  while (w.gen == g) lane_yield();
  w.progress = true;
}
template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle of <= 8 bytes"); std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
template <class T> inline T shfl_idx(T v, int src) {
  Warp &w = warp();
  const unsigned g = w.gen;
  if (src < 0 || src > 31 || w.done[src]) src = w.cur;  // (decided on arrival: the partner may leave the kernel right after)
  warp_collect(w, to_bits(v));
  return from_bits<T>(w.buf[g & 1][src]);
}
inline bool vote_all(bool p) {
  Warp &w = warp();
  const unsigned g = w.gen;
  unsigned part = 0;  // the lanes taking part, as of this lane's arrival
  for (int l = 0; l < 32; l++)
    if (!w.done[l]) part |= 1u << l;
  warp_collect(w, p ? 1u : 0u);
  for (int l = 0; l < 32; l++)
    if (((part >> l) & 1u) && !w.buf[g & 1][l]) return false;
  return true;
}
inline unsigned vote_ballot(bool p) {
  Warp &w = warp();
  const unsigned g = w.gen;
  unsigned part = 0;
  for (int l = 0; l < 32; l++)
    if (!w.done[l]) part |= 1u << l;
  warp_collect(w, p ? 1u : 0u);
  unsigned b = 0;
  for (int l = 0; l < 32; l++)
    if (((part >> l) & 1u) && w.buf[g & 1][l]) b |= 1u << l;
  return b;
}
inline void syncwarp() { warp_collect(warp(), 0); }
inline void syncthreads() {
  Warp &w = warp();
  Cta &c = *w.cta;
  const unsigned g = w.sync_gen;
  if (++w.sync_arrived == w.active) {  // last lane of the warp: meet the other warps
    w.sync_arrived = 0;
    std::unique_lock<std::mutex> lk(c.mu);
    const unsigned bg = c.bar_gen;
    if (++c.bar_count == c.nwarps) {
      c.bar_count = 0;
      c.bar_gen++;
      c.cv.notify_all();
    } else {
      c.cv.wait(lk, [&] { return c.bar_gen != bg; });
    }
    w.sync_gen++;
  }
  while (w.sync_gen == g) lane_yield();
  w.progress = true;
}

// ---- mbarriers (shared-memory objects; state kept in a side table) ----
inline void mbar_complete_if_done(MBar &b) {
  if (b.pending == 0 && b.tx == 0) {
    b.pending = (int)b.expected;
    b.phase.store(b.phase.load(std::memory_order_relaxed) ^ 1u, std::memory_order_release);
  }
}
inline void mbar_init(unsigned bar, unsigned count) {
  Cta &c = *warp().cta;
  std::lock_guard<std::mutex> lk(c.mu);
  MBar &b = c.mbars[bar];
  b.expected = count;
  b.pending = (int)count;
  b.tx = 0;
  b.phase.store(0);
}
inline void mbar_arrive_tx(unsigned bar, long bytes, bool arrive) {
  Cta &c = *warp().cta;
  std::lock_guard<std::mutex> lk(c.mu);
  auto it = c.mbars.find(bar);
  if (it == c.mbars.end()) not_emulated("an mbarrier that was never initialised");
  MBar &b = it->second;
  b.tx += bytes;
  if (arrive) {
    if (b.pending <= 0) not_emulated("more arrivals than the mbarrier expects");
    b.pending--;
  }
  mbar_complete_if_done(b);
}
inline void mbar_expect_tx(unsigned bar, unsigned bytes) { mbar_arrive_tx(bar, (long)bytes, true); }
inline void mbar_arrive(unsigned bar) { mbar_arrive_tx(bar, 0, true); }
inline bool mbar_test(unsigned bar, unsigned parity) {
  Cta &c = *warp().cta;  // (no lock: barriers are only created before the CTA-wide barrier that follows their initialisation)
  auto it = c.mbars.find(bar);
  if (it == c.mbars.end()) not_emulated("waiting on an mbarrier that was never initialised");
  return it->second.phase.load(std::memory_order_acquire) != (parity & 1u);  // the phase of that parity has completed
}
inline void mbar_wait(unsigned bar, unsigned parity) {
  while (!mbar_test(bar, parity)) lane_yield();  // (the warp's scheduler gives the core away when none of its lanes can move)
  warp().progress = true;
}
inline void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  std::memcpy(smem_ptr(dst), src, bytes);
  mbar_arrive_tx(bar, -(long)bytes, false);
}

// Shared -> global bulk copies (cp.async.bulk.global.shared::cta with bulk groups), for store-side variants of the fused
// step: the copy happens at once, so commit / wait are no-ops here -- what the emulation checks is WHAT is stored WHERE,
// not when the staging buffer may be reused (on hardware: after cp.async.bulk.wait_group.read).
inline void bulk_s2g(void *dst, unsigned src, unsigned bytes) { std::memcpy(dst, smem_ptr(src), bytes); }
inline void bulk_commit_group() {}
inline void bulk_wait_group_read(int) {}

// The thickness ring of the fused step is read one column beyond a warp's own 32 by its halo lanes (lanes 0 and 31 look at
// the neighbouring column group's slot, which that warp may be rewriting: the value only ever feeds halo-lane results
// that are thrown away -- fused_kernel.cuh, "lanes 0,1,30,31 are the x halo").  Those reads go through this function so
// that a ThreadSanitizer build does not report the one race that is there by design, and reports any other.
template <class T> __attribute__((noinline, no_sanitize("thread"))) T halo_peek(const T *p) { return *p; }

// ---- launch ----
void launch_simt(dim3 grid, dim3 block, size_t shmem, std::function<void()> body);

}  // namespace emu

template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shfl_idx(v, emu::warp().cur ^ m); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) { return emu::shfl_idx(v, emu::warp().cur - d); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) { return emu::shfl_idx(v, emu::warp().cur + d); }
inline bool __all_sync(unsigned, bool p) { return emu::vote_all(p); }
inline unsigned __ballot_sync(unsigned, bool p) { return emu::vote_ballot(p); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicMax(int *p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
inline void __syncwarp() { if (emu::cur_warp) emu::syncwarp(); }
inline void __syncthreads() { emu::syncthreads(); }
// inter-warp / inter-CTA communication through global memory (the rigid-lid wavefront solver, rigid.cuh): warps are OS
// threads here, so the GCC atomics are the real thing
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl_idx(v, src); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicExch(unsigned long long *p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T> inline T __ldcg(const T *p) { return *reinterpret_cast<const volatile T *>(p); }
template <class T> inline T __ldg(const T *p) { return *p; }
inline long long __double_as_longlong(double d) { long long b; std::memcpy(&b, &d, 8); return b; }
inline double __longlong_as_double(long long b) { double d; std::memcpy(&d, &b, 8); return d; }
