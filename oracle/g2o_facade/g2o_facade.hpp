// oracle/g2o_facade/g2o_facade.hpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference's src/slam.cpp cannot be compiled here because g2o is neither vendored nor installed
// (Dockerfile.amd64:33 clones an unpinned HEAD).  This is a facade with exactly the g2o API surface
// slam.cpp / slam.hpp use (slam.hpp:26-35; slam.cpp:53-65, 416-484, 525-550, 713-732), implemented on
// top of the oracle's restated mini-g2o (slam_oracle.cpp, orc_graph_*).  With it the reference's REAL
// slam.cpp + cone.cpp compile from where they lie (oracle/build_ref_slam.sh -> oracle/_ref/ref_slam_replay),
// so everything slam.cpp itself computes -- polar -> Cartesian conversion, the association loops of
// addConesToMap / localizer, map growth, loop-closure test, the optimise trigger, the graph-builder call
// sequence -- runs as reference code and pins the oracle's restatement of those rows
// (tests/test_pinned_by_reference.py).  The Gauss-Newton arithmetic behind optimize() is still the
// restatement (with the reference's vendored Eigen LDLT when built with -DORACLE_USE_EIGEN).
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <vector>

#include <Eigen/Dense>

extern "C" {
void* orc_graph_create();
void orc_graph_destroy(void* g);
int orc_graph_add_pose(void* g, int id, double x, double y, double th);
int orc_graph_add_landmark(void* g, int id, double x, double y);
int orc_graph_add_edge_se2(void* g, int idFrom, int idTo, const double* z3, const double* info9);
int orc_graph_add_edge_se2_xy(void* g, int poseId, int lmId, const double* z2, const double* info4);
int orc_graph_set_fixed(void* g, int id, int fixed);
int orc_graph_get_vertex(void* g, int id, double* out3);
int orc_graph_optimize(void* g, int iters, double* chi2);
double orc_normalize_theta(double t);
}

namespace g2o {

template <typename T, typename... A>
std::unique_ptr<T> make_unique(A&&... a) { return std::unique_ptr<T>(new T(std::forward<A>(a)...)); }

// types/slam2d/se2.h: rotation angle + translation; the angle is normalised by inverse() and
// operator* only, never by the constructors
class SE2 {
 public:
  SE2() : th_(0), t_(0, 0) {}
  SE2(double x, double y, double theta) : th_(theta), t_(x, y) {}
  SE2(const Eigen::Vector3d& v) : th_(v[2]), t_(v[0], v[1]) {}  // NOLINT: implicit, as in g2o
  SE2 inverse() const {
    SE2 r;
    r.th_ = orc_normalize_theta(-th_);
    r.t_ = rot(r.th_, t_ * -1.0);
    return r;
  }
  SE2 operator*(const SE2& o) const {
    SE2 r(*this);
    r.t_ += rot(th_, o.t_);
    r.th_ += o.th_;
    r.th_ = orc_normalize_theta(r.th_);
    return r;
  }
  Eigen::Vector3d toVector() const { return Eigen::Vector3d(t_[0], t_[1], th_); }

 private:
  static Eigen::Vector2d rot(double a, const Eigen::Vector2d& v) {
    const double s = std::sin(a), c = std::cos(a);  // Eigen::Rotation2D evaluates both per product
    return Eigen::Vector2d(c * v[0] - s * v[1], s * v[0] + c * v[1]);
  }
  double th_;
  Eigen::Vector2d t_;
};

struct OptimizableGraph {
  struct Vertex {
    virtual ~Vertex() {}
    void setId(int id) { id_ = id; }
    int id() const { return id_; }
    void setFixed(bool f) { fixed_ = f; }
    bool fixed() const { return fixed_; }
    int id_ = -1;
    bool fixed_ = false;
  };
  struct Edge {
    virtual ~Edge() {}
    std::vector<Vertex*>& vertices() { return v_; }
    std::vector<Vertex*> v_ = std::vector<Vertex*>(2, nullptr);
  };
};

class VertexSE2 : public OptimizableGraph::Vertex {
 public:
  void setEstimate(const SE2& e) { est_ = e; }
  const SE2& estimate() const { return est_; }
 private:
  SE2 est_;
};

class VertexPointXY : public OptimizableGraph::Vertex {
 public:
  void setEstimate(const Eigen::Vector2d& e) { est_ = e; }
  const Eigen::Vector2d& estimate() const { return est_; }
 private:
  Eigen::Vector2d est_ = Eigen::Vector2d::Zero();
};

class EdgeSE2 : public OptimizableGraph::Edge {
 public:
  void setMeasurement(const SE2& m) { z_ = m.toVector(); }
  void setInformation(const Eigen::Matrix3d& i) { info_ = i; }
  Eigen::Vector3d z_ = Eigen::Vector3d::Zero();
  Eigen::Matrix3d info_ = Eigen::Matrix3d::Identity();
};

class EdgeSE2PointXY : public OptimizableGraph::Edge {
 public:
  void setMeasurement(const Eigen::Vector2d& m) { z_ = m; }
  void setInformation(const Eigen::Matrix2d& i) { info_ = i; }
  Eigen::Vector2d z_ = Eigen::Vector2d::Zero();
  Eigen::Matrix2d info_ = Eigen::Matrix2d::Identity();
};

// solver stack: only constructed and handed over (slam.cpp:55-62)
template <int P, int L> struct BlockSolverTraits {};
struct Solver { virtual ~Solver() {} };
template <typename M> struct LinearSolverEigen {
  void setBlockOrdering(bool) {}  // false = scalar AMD, which is what the oracle's LDLT path uses
};
template <typename Traits> struct BlockSolver : Solver {
  typedef int PoseMatrixType;
  explicit BlockSolver(std::unique_ptr<LinearSolverEigen<PoseMatrixType>> ls) : ls_(std::move(ls)) {}
  std::unique_ptr<LinearSolverEigen<PoseMatrixType>> ls_;
};
struct OptimizationAlgorithm { virtual ~OptimizationAlgorithm() {} };
struct OptimizationAlgorithmGaussNewton : OptimizationAlgorithm {
  explicit OptimizationAlgorithmGaussNewton(std::unique_ptr<Solver> s) : s_(std::move(s)) {}
  std::unique_ptr<Solver> s_;
};

class SparseOptimizer {
 public:
  SparseOptimizer() {}
  SparseOptimizer(const SparseOptimizer&) = delete;
  SparseOptimizer& operator=(const SparseOptimizer&) = delete;
  ~SparseOptimizer() {
    for (auto& kv : verts_) delete kv.second;
    for (auto* e : edges_) delete e;
    delete alg_;
  }
  void setAlgorithm(OptimizationAlgorithm* a) { delete alg_; alg_ = a; }
  void setVerbose(bool) {}
  bool addVertex(OptimizableGraph::Vertex* v) {
    if (verts_.count(v->id())) return false;  // g2o refuses a second vertex with the same id
    verts_[v->id()] = v;
    return true;
  }
  bool addEdge(OptimizableGraph::Edge* e) {
    if (!e->vertices()[0] || !e->vertices()[1]) return false;
    edges_.push_back(e);
    return true;
  }
  OptimizableGraph::Vertex* vertex(int id) {
    auto it = verts_.find(id);
    return it == verts_.end() ? nullptr : it->second;
  }
  bool initializeOptimization() { return true; }
  // the whole graph is handed to the restated Gauss-Newton (insertion order of the edges kept, vertices
  // by ascending id like g2o's index mapping) and the estimates are written back
  int optimize(int iterations) {
    void* G = orc_graph_create();
    for (auto& kv : verts_) {
      if (auto* p = dynamic_cast<VertexSE2*>(kv.second)) {
        const Eigen::Vector3d e = p->estimate().toVector();
        orc_graph_add_pose(G, kv.first, e[0], e[1], e[2]);
      } else if (auto* l = dynamic_cast<VertexPointXY*>(kv.second)) {
        orc_graph_add_landmark(G, kv.first, l->estimate()[0], l->estimate()[1]);
      }
      if (kv.second->fixed()) orc_graph_set_fixed(G, kv.first, 1);
    }
    for (auto* e : edges_) {
      if (auto* o = dynamic_cast<EdgeSE2*>(e)) {
        double info[9];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) info[r * 3 + c] = o->info_(r, c);
        orc_graph_add_edge_se2(G, e->vertices()[0]->id(), e->vertices()[1]->id(), o->z_.data(), info);
      } else if (auto* m = dynamic_cast<EdgeSE2PointXY*>(e)) {
        double info[4] = {m->info_(0, 0), m->info_(0, 1), m->info_(1, 0), m->info_(1, 1)};
        orc_graph_add_edge_se2_xy(G, e->vertices()[0]->id(), e->vertices()[1]->id(), m->z_.data(), info);
      }
    }
    std::vector<double> chi2((size_t)iterations + 2, 0.0);
    const int done = orc_graph_optimize(G, iterations, chi2.data());
    for (auto& kv : verts_) {
      double out[3] = {0, 0, 0};
      orc_graph_get_vertex(G, kv.first, out);
      if (auto* p = dynamic_cast<VertexSE2*>(kv.second)) p->setEstimate(SE2(out[0], out[1], out[2]));
      else if (auto* l = dynamic_cast<VertexPointXY*>(kv.second)) l->setEstimate(Eigen::Vector2d(out[0], out[1]));
    }
    last_chi2_ = chi2;
    orc_graph_destroy(G);
    return done;
  }
  std::vector<double> last_chi2_;

 private:
  std::map<int, OptimizableGraph::Vertex*> verts_;
  std::vector<OptimizableGraph::Edge*> edges_;
  OptimizationAlgorithm* alg_ = nullptr;
};

}  // namespace g2o
