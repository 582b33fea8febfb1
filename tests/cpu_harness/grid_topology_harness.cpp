// CPU harness: exposes the host+device index arithmetic of
// graphnet_classifier_b200/csrc/grid_topology.h to Python (ctypes) so that it can
// be checked against the numpy oracle without a GPU.  Test code only.
#include "../../graphnet_classifier_b200/csrc/grid_topology.h"

extern "C" {

long long harness_num_edges(int H, int W, int diag) { return gnc::make_grid(H, W, diag).E; }

void harness_edges(int H, int W, int diag, long long* src, long long* dst) {
  gnc::GridDims g = gnc::make_grid(H, W, diag);
  for (int64_t e = 0; e < g.E; ++e) {
    int64_t s, d;
    gnc::grid_edge(g, e, s, d);
    src[e] = s; dst[e] = d;
  }
}

// which = 1: CSR by destination (in-edges); which = 0: by source (out-edges)
void harness_csr(int H, int W, int diag, int which, int* rowptr, int* eid) {
  gnc::GridDims g = gnc::make_grid(H, W, diag);
  for (int64_t v = 0; v < g.N; ++v) {
    int64_t ids[4], before;
    int n = which ? gnc::grid_in_edges(g, v, ids, &before) : gnc::grid_out_edges(g, v, ids, &before);
    rowptr[v] = (int)before;
    for (int k = 0; k < n; ++k) eid[before + k] = (int)ids[k];
    if (v == g.N - 1) rowptr[g.N] = (int)(before + n);
  }
}
}
