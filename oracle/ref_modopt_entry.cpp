// ref_modopt_entry.cpp -- C entry points around the reference's OWN community-detection classes
// (TEST INFRASTRUCTURE).  The reference translation unit is included from where it lies under
// $(REF)/src at build time (oracle/Makefile, target `ref`); nothing of it is copied here.
//   ref_net_build    ModularityOptimizer::matrixToNetwork          (ModularityOptimizer.cpp:761-806)
//   ref_net_quality  VOSClusteringTechnique::calcQualityFunction   (:462-482)
//   ref_net_reduce   Network::createReducedNetwork                 (:322-373)
//   ref_louvain_hooked   the driver loop of RModularityOptimizer.cpp:101-172 (== the STANDALONE main,
//                    ModularityOptimizer.cpp:935-985) and the recursions runLouvainAlgorithm :585-604 /
//                    runLouvainAlgorithmWithMultilevelRefinement :618-637, restated here so that the three
//                    bulk steps can be supplied from outside (callbacks) while the reference's own
//                    runLocalMovingAlgorithm, JavaRandom, mergeClusters and orderClustersByNNodes do the
//                    rest.  With null callbacks it is the reference's algorithm end to end.
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

// ---------------------------------------------------------------------------------------------
// Louvain with the bulk steps supplied from outside
// ---------------------------------------------------------------------------------------------
extern "C" {
// lower-triangle edge list -> CSR arrays (first[n_nodes+1], neighbor/edge_w[2m], node_w[n_nodes], *total_w)
typedef void (*hook_network_fn)(const int* node1, const int* node2, const double* w, long long m, int n_nodes,
                                int* first, int* neighbor, double* edge_w, double* node_w, double* total_w);
typedef double (*hook_quality_fn)(int n_nodes, const int* first, const int* neighbor, const double* edge_w,
                                  const double* node_w, double self_links, const int* cluster, int n_clusters,
                                  double resolution);
// returns the reduced edge count; r_neighbor / r_edge_w have room for the parent's edge count
typedef long long (*hook_reduce_fn)(int n_nodes, const int* first, const int* neighbor, const double* edge_w,
                                    const double* node_w, double self_links, const int* cluster, int n_clusters,
                                    int* r_first, int* r_neighbor, double* r_edge_w, double* r_node_w,
                                    double* r_self_links);
}

namespace {
using namespace ModularityOptimizer;

struct Hooks {
  hook_network_fn network;
  hook_quality_fn quality;
  hook_reduce_fn reduce;
  long long calls[3];
};

// a Network assembled from arrays, self links included (the member is protected)
struct Assembled : Network {
  Assembled(int n, DVector& nw, IVector& first, IVector& nb, DVector& ew, double self_links)
      : Network(n, &nw, first, nb, &ew) {
    totalEdgeWeightSelfLinks = self_links;
  }
};

std::shared_ptr<Network> hooked_reduce(Network& net, const Clustering& cl, Hooks& h) {
  if (!h.reduce) return std::make_shared<Network>(net.createReducedNetwork(cl));
  Peek p(net);
  const int nc = cl.nClusters;
  const size_t cap = p.nbr().size() ? p.nbr().size() : 1;
  IVector first(nc + 1), nb(cap);
  DVector ew(cap), nw(nc);
  double self = 0.0;
  h.calls[2]++;
  const long long e = h.reduce(net.getNNodes(), p.fni().data(), p.nbr().data(), p.ew().data(), p.nw().data(),
                               net.getTotalEdgeWeightSelfLinks(), cl.cluster.data(), nc, first.data(), nb.data(),
                               ew.data(), nw.data(), &self);
  nb.resize(e);
  ew.resize(e);
  return std::make_shared<Network>(Assembled(nc, nw, first, nb, ew, self));
}

double hooked_quality(VOSClusteringTechnique& vos, Hooks& h) {
  if (!h.quality) return vos.calcQualityFunction();
  Peek p(*vos.getNetwork());
  const Clustering& cl = *vos.getClustering();
  h.calls[1]++;
  return h.quality(p.getNNodes(), p.fni().data(), p.nbr().data(), p.ew().data(), p.nw().data(),
                   p.getTotalEdgeWeightSelfLinks(), cl.cluster.data(), cl.nClusters, vos.getResolution());
}

// :585-604 (refine == false) and :618-637 (refine == true)
bool hooked_louvain(VOSClusteringTechnique& vos, JavaRandom& random, bool refine, Hooks& h) {
  if (vos.getNetwork()->getNNodes() == 1) return false;
  bool update = vos.runLocalMovingAlgorithm(random);
  std::shared_ptr<Clustering> cl = vos.getClustering();
  if (cl->nClusters < vos.getNetwork()->getNNodes()) {
    VOSClusteringTechnique vos2(hooked_reduce(*vos.getNetwork(), *cl, h), vos.getResolution());
    const bool update2 = hooked_louvain(vos2, random, refine, h);
    if (update2) {
      update = true;
      cl->mergeClusters(*vos2.getClustering());
      if (refine) vos.runLocalMovingAlgorithm(random);
    }
  }
  return update;
}
}  // namespace

extern "C" {

// labels[n_nodes] (after orderClustersByNNodes), *max_modularity, calls[3] = how often each hook ran.
// Returns the number of nodes, or -1 on an exception.
int ref_louvain_hooked(const int* node1, const int* node2, const double* w, long long m, double resolution,
                       int algorithm, int n_random_starts, int n_iterations, unsigned long long seed,
                       hook_network_fn hn, hook_quality_fn hq, hook_reduce_fn hr, int* labels,
                       double* max_modularity, long long* calls) {
  try {
    Hooks h = {hn, hq, hr, {0, 0, 0}};
    std::shared_ptr<Network> network;
    double total_w;
    if (hn) {
      int n_nodes = 0;
      for (long long e = 0; e < m; ++e) n_nodes = std::max(n_nodes, std::max(node1[e], node2[e]) + 1);
      IVector first(n_nodes + 1), nb(2 * m);
      DVector ew(2 * m), nw(n_nodes);
      h.calls[0]++;
      hn(node1, node2, w, m, n_nodes, first.data(), nb.data(), ew.data(), nw.data(), &total_w);
      network = std::make_shared<Network>(Assembled(n_nodes, nw, first, nb, ew, 0.0));
    } else {
      IVector a(node1, node1 + m), b(node2, node2 + m);
      DVector ww(w, w + m);
      network = matrixToNetwork(a, b, ww, 1);
      total_w = network->getTotalEdgeWeight();
    }
    // RModularityOptimizer.cpp:101 (modularity function 1)
    const double resolution2 = resolution / (2 * total_w + network->getTotalEdgeWeightSelfLinks());
    std::shared_ptr<Clustering> best;
    double max_mod = -std::numeric_limits<double>::infinity(), modularity = 0.0;
    JavaRandom random(seed);
    for (int i = 0; i < n_random_starts; i++) {
      VOSClusteringTechnique vos(network, resolution2);
      int j = 0;
      bool update = true;
      do {
        update = hooked_louvain(vos, random, algorithm == 2, h);
        j++;
        modularity = hooked_quality(vos, h);
      } while ((j < n_iterations) && update);
      if (modularity > max_mod) {
        best = vos.getClustering();
        max_mod = modularity;
      }
    }
    best->orderClustersByNNodes();
    std::copy(best->cluster.begin(), best->cluster.end(), labels);
    *max_modularity = max_mod;
    for (int q = 0; q < 3; ++q) calls[q] = h.calls[q];
    return network->getNNodes();
  } catch (...) {
    return -1;
  }
}

}  // extern "C"
