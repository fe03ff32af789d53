// ref_modopt_entry.cpp -- C entry points around the reference's OWN community-detection classes
// (TEST INFRASTRUCTURE).  The reference translation unit is included from where it lies under
// $(REF)/src at build time (oracle/Makefile, target `ref`); nothing of it is copied here.
//   ref_net_build    ModularityOptimizer::matrixToNetwork          (ModularityOptimizer.cpp:761-806)
//   ref_net_quality  VOSClusteringTechnique::calcQualityFunction   (:462-482)
//   ref_net_reduce   Network::createReducedNetwork                 (:322-373)
#include "ModularityOptimizer.cpp"

namespace {
// the network's arrays are protected members: a derived class may read them
struct Peek : ModularityOptimizer::Network {
  explicit Peek(const ModularityOptimizer::Network& n) : ModularityOptimizer::Network(n) {}
  const IVector& fni() const { return firstNeighborIndex; }
  const IVector& nbr() const { return neighbor; }
  const DVector& ew() const { return edgeWeight; }
  const DVector& nw() const { return nodeWeight; }
  int directed_edges() const { return nEdges; }
};
typedef std::shared_ptr<ModularityOptimizer::Network> NetPtr;
}  // namespace

extern "C" {

void* ref_net_build(const int* node1, const int* node2, const double* w, long long m, int modularity_function) {
  IVector a(node1, node1 + m), b(node2, node2 + m);
  DVector ww(w, w + m);
  return new NetPtr(ModularityOptimizer::matrixToNetwork(a, b, ww, modularity_function));
}

void ref_net_dims(void* h, int* n_nodes, int* n_directed_edges) {
  Peek p(**(NetPtr*)h);
  *n_nodes = p.getNNodes();
  *n_directed_edges = p.directed_edges();
}

void ref_net_get(void* h, int* first, int* neighbor, double* edge_w, double* node_w, double* total_w,
                 double* self_links) {
  Peek p(**(NetPtr*)h);
  std::copy(p.fni().begin(), p.fni().end(), first);
  std::copy(p.nbr().begin(), p.nbr().end(), neighbor);
  std::copy(p.ew().begin(), p.ew().end(), edge_w);
  std::copy(p.nw().begin(), p.nw().end(), node_w);
  *total_w = p.getTotalEdgeWeight();
  *self_links = p.getTotalEdgeWeightSelfLinks();
}

double ref_net_quality(void* h, const int* cluster, double resolution) {
  NetPtr net = *(NetPtr*)h;
  IVector cl(cluster, cluster + net->getNNodes());
  auto clustering = std::make_shared<ModularityOptimizer::Clustering>(cl);
  ModularityOptimizer::VOSClusteringTechnique vos(net, clustering, resolution);
  return vos.calcQualityFunction();
}

void* ref_net_reduce(void* h, const int* cluster) {
  NetPtr net = *(NetPtr*)h;
  IVector cl(cluster, cluster + net->getNNodes());
  ModularityOptimizer::Clustering clustering(cl);
  return new NetPtr(std::make_shared<ModularityOptimizer::Network>(net->createReducedNetwork(clustering)));
}

void ref_net_free(void* h) { delete (NetPtr*)h; }

}  // extern "C"
