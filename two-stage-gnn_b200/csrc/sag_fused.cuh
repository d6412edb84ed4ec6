// sag_fused.cuh -- interface between the step executor (k10_sag_exec.cu, owner of the arena layout) and the
// graph-resident SAGPool forward (k13_sag_fused.cu).
#pragma once
#include "common.cuh"

namespace tsg {

struct FusedArgs {
  // batch (compact form: labels + graph-local endpoints, lists coalesced and symmetric -- verified per graph)
  int G, L, H;                       // graphs, node-label classes (= conv1 in_channels), hidden width
  int nmax[3];                       // max nodes of one graph at pooling level 0, 1, 2
  int emax;                          // max directed edges of one graph
  float ratio; double avg_degree;    // pooling ratio (caps of the size classes), directed edges per node of the batch
  int cls, cls_nlo, cls_elo;         // set per launch: size class index; graphs with n0 <= cls_nlo && e <= cls_elo are not ours
  const int32_t* label; const int32_t* lrow; const int32_t* lcol;
  const int64_t* edge_ptr;           // [G+1]
  const int64_t* level_ptr;          // [4, G+1]
  const float* params[12];           // per level: conv W [in,H], conv b [H], score w [H,1], score b [1]
  // saved tensors (arena)
  float* h[3]; float* score[3]; int64_t* perm[3]; int32_t* argmax[3];
  float* z;                          // [G, 2H]
  // bookkeeping
  unsigned* sched;                   // 8 zeroed counters: dynamic graph scheduling, one per size class / direction
  int* status;                       // OR of TSG_FUSED_* bits
};

enum { TSG_FUSED_BAD_EDGE = 1, TSG_FUSED_UNSORTED = 2, TSG_FUSED_ASYMMETRIC = 4 };

bool fused_supported(int H, int L, int nmax0, int nmax1, int emax);     // fits shared memory / index widths
int fused_num_ctas();
int launch_sag_fused_fwd(const FusedArgs& a, cudaStream_t st);

}  // namespace tsg
