// ORACLE (test infrastructure, NOT product code) -- C wrapper around the REFERENCE's own max-clique finder, compiled from the
// sources where they lie under /root/reference (swarm_localization/src/swarm_outlier_rejection/third_party/
// fast_max-clique_finder/src: findCliqueHeu.cpp, findClique.cpp, graphIO.cpp, utils.cpp -- plain C++, no external library).
// Built by oracle/ref_build/Makefile into oracle/_ref/libfmc_ref.so; nothing of the reference is copied into this repository.
//
// The graph is handed over exactly as OutlierRejectionLoopEdgesPCM does (swarm_outlier_rejection.cpp:277-287): CSR offsets in
// m_vi_Vertices, neighbour lists in m_vi_Edges, CalculateVertexDegrees(), FMC::maxCliqueHeu(gio, data).
#include "findClique.h"

extern "C" int fmc_max_clique_heu(int n, const int* offsets /*[n+1]*/, const int* edges, int* clique_out /*[n]*/) {
  FMC::CGraphIO gio;
  gio.m_vi_Vertices.assign(offsets, offsets + n + 1);
  gio.m_vi_Edges.assign(edges, edges + offsets[n]);
  gio.CalculateVertexDegrees();
  std::vector<int> data;
  const int size = FMC::maxCliqueHeu(gio, data);
  for (size_t i = 0; i < data.size() && (int)i < n; ++i) clique_out[i] = data[i];
  return (int)data.size() == size ? size : -(int)data.size() - 1000000;   // (the two always agree; a mismatch is reported)
}
