/* modopt_oracle.c -- CPU restatement of the data-parallel pieces of the reference's community
 * detection (TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and
 * bench.py's checker legs may load it).
 *
 * Follows, in the reference's own (sequential) order of operations so that every double is the
 * reference's double:
 *   modopt_network   matrixToNetwork, src/ModularityOptimizer.cpp:761-806, with the Network
 *                    constructor :169-188 (node weights :272-284) and getTotalEdgeWeight :268-270
 *   modopt_quality   VOSClusteringTechnique::calcQualityFunction :462-482
 *   modopt_reduce    Network::createReducedNetwork :322-373 (nodes per cluster :106-118)
 * Pinned against the reference's own classes compiled unmodified (oracle/ref_modopt_entry.cpp ->
 * oracle/_ref/libgficf_ref_modopt.so) by tests/test_network_oracle.py, and against the fixtures
 * tests/golden/net_*.npz that were generated through those classes. */
#include <stdlib.h>
#include <string.h>

/* edge list (node1[e] < node2[e] are kept, like :768) -> CSR; arrays sized by the caller:
 * first[n_nodes+1], neighbor/edge_w[2*m], node_w[n_nodes].  Returns the number of directed edges. */
long long modopt_network(const int* node1, const int* node2, const double* w, long long m, int n_nodes,
                         int* first, int* neighbor, double* edge_w, double* node_w, double* total_w) {
  int* deg = (int*)calloc((size_t)n_nodes + 1, sizeof(int));
  long long e, n_edges = 0;
  int v;
  for (e = 0; e < m; ++e)
    if (node1[e] < node2[e]) {
      deg[node1[e]]++;
      deg[node2[e]]++;
    }
  for (v = 0; v < n_nodes; ++v) {
    first[v] = (int)n_edges;
    n_edges += deg[v];
  }
  first[n_nodes] = (int)n_edges;
  memset(deg, 0, ((size_t)n_nodes + 1) * sizeof(int));
  for (e = 0; e < m; ++e)
    if (node1[e] < node2[e]) {
      int a = node1[e], b = node2[e];
      long long j = (long long)first[a] + deg[a]++;
      neighbor[j] = b;
      edge_w[j] = w[e];
      j = (long long)first[b] + deg[b]++;
      neighbor[j] = a;
      edge_w[j] = w[e];
    }
  free(deg);
  for (v = 0; v < n_nodes; ++v) { /* std::accumulate(first, last, 0.0) per node */
    double s = 0.0;
    long long k;
    for (k = first[v]; k < first[v + 1]; ++k) s += edge_w[k];
    node_w[v] = s;
  }
  {
    double s = 0.0;
    for (e = 0; e < n_edges; ++e) s += edge_w[e];
    *total_w = s / 2.0;
  }
  return n_edges;
}

/* std::accumulate(x, x + n, s0): what Network::getTotalEdgeWeight (:268-270) does before halving */
double modopt_seq_sum(const double* x, long long n, double s0) {
  long long i;
  for (i = 0; i < n; ++i) s0 += x[i];
  return s0;
}

/* cluster_w[n_clusters] is an output as well (the weights :474-476 forms on the way) */
double modopt_quality(int n_nodes, const int* first, const int* neighbor, const double* edge_w,
                      const double* node_w, double self_links, double total_w, const int* cluster,
                      int n_clusters, double resolution, double* cluster_w) {
  double q = 0.0;
  int i, k;
  for (i = 0; i < n_nodes; ++i) {
    const int j = cluster[i];
    for (k = first[i]; k < first[i + 1]; ++k)
      if (cluster[neighbor[k]] == j) q += edge_w[k];
  }
  q += self_links;
  for (i = 0; i < n_clusters; ++i) cluster_w[i] = 0.0;
  for (i = 0; i < n_nodes; ++i) cluster_w[cluster[i]] += node_w[i];
  for (i = 0; i < n_clusters; ++i) q -= cluster_w[i] * cluster_w[i] * resolution;
  q /= 2 * total_w + self_links;
  return q;
}

/* r_first[n_clusters+1], r_neighbor / r_edge_w sized for the parent's edge count, r_node_w[n_clusters];
 * *self_links: in = the parent's total, out = the reduced network's.  Returns the reduced edge count. */
long long modopt_reduce(int n_nodes, const int* first, const int* neighbor, const double* edge_w,
                        const double* node_w, const int* cluster, int n_clusters, int* r_first,
                        int* r_neighbor, double* r_edge_w, double* r_node_w, double* self_links) {
  /* nodes per cluster, ascending node id inside a cluster */
  int* cstart = (int*)calloc((size_t)n_clusters + 1, sizeof(int));
  int* fill = (int*)calloc((size_t)n_clusters + 1, sizeof(int));
  int* nodes = (int*)malloc(((size_t)n_nodes + 1) * sizeof(int));
  int* seen_list = (int*)malloc(((size_t)n_clusters + 1) * sizeof(int));
  double* acc = (double*)calloc((size_t)n_clusters + 1, sizeof(double));
  long long n_red = 0;
  double self = *self_links;
  int i, c;
  for (i = 0; i < n_nodes; ++i) cstart[cluster[i] + 1]++;
  for (c = 0; c < n_clusters; ++c) cstart[c + 1] += cstart[c];
  for (i = 0; i < n_nodes; ++i) nodes[cstart[cluster[i]] + fill[cluster[i]]++] = i;
  r_first[0] = 0;
  for (c = 0; c < n_clusters; ++c) {
    int n_seen = 0, t, k;
    r_node_w[c] = 0.0;
    for (t = cstart[c]; t < cstart[c + 1]; ++t) {
      const int l = nodes[t];
      r_node_w[c] += node_w[l];
      for (k = first[l]; k < first[l + 1]; ++k) {
        const int n = cluster[neighbor[k]];
        if (n != c) {
          if (acc[n] == 0) seen_list[n_seen++] = n; /* :342: "no weight yet" means first appearance */
          acc[n] += edge_w[k];
        } else {
          self += edge_w[k];
        }
      }
    }
    for (k = 0; k < n_seen; ++k) {
      r_neighbor[n_red + k] = seen_list[k];
      r_edge_w[n_red + k] = acc[seen_list[k]];
      acc[seen_list[k]] = 0;
    }
    n_red += n_seen;
    r_first[c + 1] = (int)n_red;
  }
  *self_links = self;
  free(cstart);
  free(fill);
  free(nodes);
  free(seen_list);
  free(acc);
  return n_red;
}
