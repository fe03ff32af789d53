// network_emu.cpp -- the network kernels and their launch sequences (gficf_b200/csrc/network_plan.h)
// compiled as plain C++ against the CUDA emulation (TEST INFRASTRUCTURE).  Built and loaded by
// tests/test_network_emu.py:
//   g++ -O1 -std=c++17 -ffp-contract=off -DGFICF_CUDA_EMU -Itests/cuda_emu -Igficf_b200/csrc -Iinclude -shared -fPIC
#include "network_plan.h"

#include <vector>

using namespace gficf;

namespace {
struct Ctx {
  std::vector<char> scratch;
  unsigned flags = 0;
  NetCtx cx;
  ~Ctx() {
#ifdef GFICF_NET_ASAN
    __asan_unpoison_memory_region(scratch.data(), scratch.size());
#endif
  }
  Ctx(long long nn, long long cap, int max_ctas) {
    scratch.assign(net_scratch_layout(nullptr, nn, cap).bytes, (char)0xA5);  // scratch is never assumed zero
    cx.st = 0;
#ifdef GFICF_NET_ASAN
    __asan_poison_memory_region(scratch.data(), scratch.size());  // the layout makes the pieces addressable again
#endif
    cx.sc = net_scratch_layout(scratch.data(), nn, cap);
    cx.flags = &flags;
    cx.max_ctas = max_ctas;
  }
};
}  // namespace

extern "C" {

long long emu_launch_count() { return cuda_emu::g_launches; }

// exclusive scan of int counts: out[m+1]
void emu_scan(const int* cnt, long long m, long long* out) {
  Ctx c(1, m, 3);
  net_scan(c.cx, cnt, m, out);
}

// stable sort of (keys, vals) by the low nbits; results copied back into keys / vals
void emu_radix_sort(unsigned long long* keys, unsigned* vals, long long n, int nbits, int max_ctas) {
  Ctx c(1, n, max_ctas);
  memcpy(c.cx.sc.keys[0], keys, (size_t)n * 8);
  memcpy(c.cx.sc.vals[0], vals, (size_t)n * 4);
  const int b = net_radix_sort(c.cx, 0, n, nbits);
  memcpy(keys, c.cx.sc.keys[b], (size_t)n * 8);
  memcpy(vals, c.cx.sc.vals[b], (size_t)n * 4);
}

// ((s0 + x[0]) + x[1]) + ... replayed exactly, times scale
double emu_seq_sum(const double* x, long long n, double s0, double scale, unsigned* flags, int max_ctas) {
  Ctx c(1, 1, max_ctas);
  double out = -1.0;
  net_seq_sum(c.cx, x, n, s0, scale, &out);
  *flags = c.flags;
  return out;
}

unsigned emu_net_build(const long long* colptr, const int* row, const double* w, long long nv, long long nnz,
                       long long* first, int* neighbor, double* edge_w, double* node_w, double* total_w,
                       int max_ctas) {
  Ctx c(nv, nnz, max_ctas);
  net_build(c.cx, colptr, row, w, nv, nnz, first, neighbor, edge_w, node_w, total_w);
  return c.flags;
}

unsigned emu_net_quality(const long long* first, const int* neighbor, const double* edge_w, const double* node_w,
                         long long n_nodes, const int* cluster, int n_clusters, double resolution,
                         double self_links, double total_w, double* cluster_w, double* q, int max_ctas) {
  Ctx c(n_nodes, first[n_nodes], max_ctas);
  net_quality(c.cx, first, neighbor, edge_w, node_w, n_nodes, first[n_nodes], cluster, n_clusters, resolution,
              self_links, &total_w, cluster_w, q);
  return c.flags;
}

long long emu_net_reduce(const long long* first, const int* neighbor, const double* edge_w, const double* node_w,
                         long long n_nodes, const int* cluster, int n_clusters, double self_links,
                         long long* r_first, int* r_neighbor, double* r_edge_w, long long r_cap, double* r_node_w,
                         double* r_self_links, double* r_total_w, long long* n_needed, unsigned* flags,
                         int max_ctas) {
  Ctx c(n_nodes, first[n_nodes], max_ctas);
  const long long r = net_reduce(c.cx, first, neighbor, edge_w, node_w, n_nodes, first[n_nodes], cluster, n_clusters,
                                 self_links, r_first, r_neighbor, r_edge_w, r_cap, r_node_w, r_self_links,
                                 r_total_w, n_needed);
  *flags = c.flags;
  return r;
}


// ---- the C entry points themselves, same names and signatures as include/gficf_cuda.h, with host
//      memory standing in for device memory: lets the CPU suite drive gficf_b200/modularity.py (the
//      Python mirror) through every argument it passes.
size_t gficf_cuda_network_scratch_bytes(int64_t n_nodes, int64_t n_items) {
  if (n_nodes < 0 || n_items < 0) return 0;
  return net_scratch_layout(nullptr, n_nodes, n_items).bytes;
}

int gficf_cuda_network_dev(const int64_t* d_colptr, const int32_t* d_row, const double* d_w, int64_t n_vertices,
                           int64_t nnz, int64_t* d_first, int32_t* d_neighbor, double* d_edge_w, double* d_node_w,
                           double* d_total_w, void* d_scratch, size_t scratch_bytes, uint32_t* d_flags, void*) {
  return net_entry_network(d_colptr, d_row, d_w, n_vertices, nnz, d_first, d_neighbor, d_edge_w, d_node_w, d_total_w,
                           d_scratch, scratch_bytes, d_flags, 0, 2);
}

int gficf_cuda_network_quality_dev(const int64_t* d_first, const int32_t* d_neighbor, const double* d_edge_w,
                                   const double* d_node_w, int64_t n_nodes, int64_t n_edges,
                                   const int32_t* d_cluster, int32_t n_clusters, double resolution,
                                   double self_links, const double* d_total_w, double* d_cluster_w,
                                   double* d_quality, void* d_scratch, size_t scratch_bytes, uint32_t* d_flags,
                                   void*) {
  return net_entry_quality(d_first, d_neighbor, d_edge_w, d_node_w, n_nodes, n_edges, d_cluster, n_clusters,
                           resolution, self_links, d_total_w, d_cluster_w, d_quality, d_scratch, scratch_bytes,
                           d_flags, 0, 2);
}

int gficf_cuda_network_reduce_dev(const int64_t* d_first, const int32_t* d_neighbor, const double* d_edge_w,
                                  const double* d_node_w, int64_t n_nodes, int64_t n_edges,
                                  const int32_t* d_cluster, int32_t n_clusters, double self_links,
                                  int64_t* d_r_first, int32_t* d_r_neighbor, double* d_r_edge_w, int64_t r_cap,
                                  double* d_r_node_w, double* d_r_self_links, double* d_r_total_w,
                                  int64_t* n_reduced_edges, void* d_scratch, size_t scratch_bytes,
                                  uint32_t* d_flags, void*) {
  return net_entry_reduce(d_first, d_neighbor, d_edge_w, d_node_w, n_nodes, n_edges, d_cluster, n_clusters, self_links,
                          d_r_first, d_r_neighbor, d_r_edge_w, r_cap, d_r_node_w, d_r_self_links, d_r_total_w,
                          n_reduced_edges, d_scratch, scratch_bytes, d_flags, 0, 2);
}

}  // extern "C"
