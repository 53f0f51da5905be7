// Persistent work-queue body shared by the binary PC-SAFT and group-contribution bubble / dew kernels.
//
// Every lane runs the resumable PointMachine (mix_machine.cuh) one Helmholtz evaluation per trip; lanes
// whose state has finished take the next state from a global queue (warp-aggregated atomic, batched so
// that the divergent set-up -- parameter load + coefficient stage -- is amortised over >= kPointRefillMin
// lanes) instead of idling until the slowest lane of the warp is done.  The evaluation is entered by all
// 32 lanes together; only the cheap bookkeeping in `advance` diverges by solver phase.
// `Source` provides
//   void dummy(Model&)                                  a harmless model for lanes that hold no state yet
//   bool load(int idx, Model&, double& z0, double& T)   set up state idx; false = invalid composition
//   void store(int idx, int rc, bool bubble, double z0, double T, const PointMachine<Model>&)
#pragma once
#include "mix_machine.cuh"

namespace feos {

constexpr int kPointRefillMin = 8;

template <class Model, class Source>
__device__ __forceinline__ void point_queue_body(const Source& src, Model& mdl, int n, bool bubble,
                                                 int32_t* __restrict__ evals, int32_t* __restrict__ retry_list,
                                                 int32_t* __restrict__ retry_count, int32_t* __restrict__ queue_head) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  bool active = false, exhausted = false;
  int done = kRunning;  // status of a finished state whose results are still to be written
  int idx = 0, ne = 0;
  double T = 0.0, z0 = 0.5;
  src.dummy(mdl);
  PointMachine<Model> m;
  m.start(mdl, bubble, 0.5);
  while (true) {
    unsigned idle = __ballot_sync(full, !active);
    if (idle) {
      int nidle = __popc(idle);
      if ((!exhausted && nidle >= kPointRefillMin) || nidle == 32) {
        if (done != kRunning) {
          if (evals) evals[idx] = ne;
          if (done == kStatusRetry) retry_list[atomicAdd(retry_count, 1)] = idx;
          else src.store(idx, done, bubble, z0, T, m);
          done = kRunning;
        }
        if (!exhausted) {
          int base = 0;
          if (lane == 0) base = atomicAdd(queue_head, nidle);
          base = __shfl_sync(full, base, 0);
          int mine = base + __popc(idle & ((1u << lane) - 1u));
          if (!active && mine < n) {
            idx = mine;
            ne = 0;
            if (src.load(idx, mdl, z0, T)) {
              m.start(mdl, bubble, z0);
              active = true;
            } else {
              done = kStatusFailed;
            }
          }
          if (base + nidle >= n) exhausted = true;
          idle = __ballot_sync(full, !active);
        }
      }
      if (exhausted && idle == full) {
        if (done != kRunning) {  // what finished since the last refill (reached with every lane idle)
          if (evals) evals[idx] = ne;
          if (done == kStatusRetry) retry_list[atomicAdd(retry_count, 1)] = idx;
          else src.store(idx, done, bubble, z0, T, m);
        }
        break;
      }
    }
    double r0, r1;
    m.density(r0, r1);
    __syncwarp();
    MixEval e = mdl.eval(r0, r1);
    if (active) {
      ++ne;
      int st = m.advance(mdl, e);
      if (st != kRunning) {
        active = false;
        done = st;
      }
    }
  }
}

}  // namespace feos
