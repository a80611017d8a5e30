// Mesh simplification of ConvertToBinary (source/mesh_stream/ConvertToBinary.cpp:186-203): quadric-error edge contraction
// after Garland & Heckbert as source/render/MeshSimplifier.cpp runs it — sweeps over the face list in order, contracting
// every edge whose cost is under a percentile threshold, with immediate in-place updates of the neighbourhood.  Each
// contraction reads the state the previous one left (vertex positions, quadrics, face costs, touched / deleted flags), so
// the reference's result is defined by that sequential order and the stage is host code here exactly as it is there
// (kThreads = 1 in the reference's call); the GPU delivers the mesh it starts from (derp_mesh.cuh) in double precision.
//
// Arithmetic conventions (they decide threshold comparisons, hence the output): IEEE double, no FMA contraction,
// 3-term sums left to right, cross product and 3 x 3 determinant in the textbook cofactor order Eigen's fixed-size
// kernels use.  Checked against the reference's own MeshSimplifier.cpp compiled into oracle/_ref (tests/test_mesh.py).
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <set>
#include <vector>

namespace derp {
namespace simplify {

struct V3 {
  double x, y, z;
};
inline V3 sub(const V3& a, const V3& b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline double dot(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(const V3& a, const V3& b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 unit(const V3& a) {
  const double n = std::sqrt(dot(a, a));
  return n > 0 ? V3{a.x / n, a.y / n, a.z / n} : a;
}

struct Quadric {  // symmetric 4 x 4, stored in full like the reference's Matrix4d (sums are element-wise either way)
  double m[4][4];
};
inline void addInto(Quadric& a, const Quadric& b) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) a.m[i][j] = a.m[i][j] + b.m[i][j];
}
inline double det3(const double a[3][3]) {
  auto term = [&](int i, int j, int k) { return a[0][i] * (a[1][j] * a[2][k] - a[1][k] * a[2][j]); };
  return term(0, 1, 2) - term(1, 0, 2) + term(2, 0, 1);
}
// v^T Q v for homogeneous v = (x, y, z, 1), symmetric Q (MeshSimplifier.cpp computeFastError: same term order)
inline double quadricError(const Quadric& q, const V3& v) {
  return q.m[0][0] * v.x * v.x + 2 * q.m[0][1] * v.x * v.y + 2 * q.m[0][2] * v.x * v.z + 2 * q.m[0][3] * v.x +
      q.m[1][1] * v.y * v.y + 2 * q.m[1][2] * v.y * v.z + 2 * q.m[1][3] * v.y + q.m[2][2] * v.z * v.z +
      2 * q.m[2][3] * v.z + q.m[3][3];
}

class Mesh {
 public:
  struct Vertex {
    std::vector<int> faces;
    V3 p{0, 0, 0};
    Quadric q{};
    bool boundary = false;
  };
  struct Face {  // (the plane quadric of a face is only needed to seed the vertex quadrics: it is not kept here,
                 // which makes the per-sweep compaction of the face list three times lighter than the reference's)
    int v[3];
    V3 normal{0, 0, 0};
  };
  std::vector<Vertex> verts;
  std::vector<Face> faces;
  // what the sweeps scan, apart from the face records: 3 edge costs per face (edge j = v[j] -> v[j + 1]) and the flags
  std::vector<double> cost;
  std::vector<uint8_t> flag;
  enum : uint8_t { kDeleted = 1, kTouched = 2 };

  // xyz: 3 doubles per vertex; idx: 3 indices per face
  Mesh(const double* xyz, size_t nv, const uint32_t* idx, size_t nf) : verts(nv), faces(nf), cost(3 * nf), flag(nf, 0) {
    for (size_t i = 0; i < nv; ++i) verts[i].p = V3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    for (size_t i = 0; i < nf; ++i)
      for (int j = 0; j < 3; ++j) faces[i].v[j] = (int)idx[3 * i + j];
  }

  // optimal position of the vertex an edge contracts to, and the error there (computeError, equi-error variant)
  double contraction(const Vertex& a, const Vertex& b, V3* target) const {
    Quadric q = a.q;
    addInto(q, b.q);
    double top[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) top[i][j] = q.m[i][j];
    const double det = det3(top);
    if (det != 0 && !(a.boundary && b.boundary)) {
      // Cramer: the first three entries of the last column of the inverse of [top | q14 q24 q34; 0 0 0 1]
      const double mx[3][3] = {{q.m[0][1], q.m[0][2], q.m[0][3]}, {q.m[1][1], q.m[1][2], q.m[1][3]}, {q.m[2][1], q.m[2][2], q.m[2][3]}};
      const double my[3][3] = {{q.m[0][0], q.m[0][2], q.m[0][3]}, {q.m[1][0], q.m[1][2], q.m[1][3]}, {q.m[2][0], q.m[2][2], q.m[2][3]}};
      const double mz[3][3] = {{q.m[0][0], q.m[0][1], q.m[0][3]}, {q.m[1][0], q.m[1][1], q.m[1][3]}, {q.m[2][0], q.m[2][1], q.m[2][3]}};
      const double inv = 1 / det;
      *target = V3{(-det3(mx)) * inv, det3(my) * inv, (-det3(mz)) * inv};
      return quadricError(q, *target);
    }
    const V3 cand[3] = {a.p, b.p, V3{(a.p.x + b.p.x) / 2, (a.p.y + b.p.y) / 2, (a.p.z + b.p.z) / 2}};
    int best = 0;
    double err[3];
    for (int k = 0; k < 3; ++k) {
      err[k] = quadricError(q, cand[k]);
      if (err[k] < err[best]) best = k;  // std::min_element: first of equal minima
    }
    *target = cand[best];
    return err[best];
  }

  void initialQuadrics() {
    for (Face& f : faces) {  // face order = accumulation order of the vertex quadrics, as in the reference
      const V3 &p0 = verts[f.v[0]].p, &p1 = verts[f.v[1]].p, &p2 = verts[f.v[2]].p;
      const V3 n = unit(cross(sub(p1, p0), sub(p2, p0)));
      f.normal = n;
      const double plane[4] = {n.x, n.y, n.z, -dot(n, p0)};
      Quadric q;
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) q.m[i][j] = plane[i] * plane[j];
      for (int j = 0; j < 3; ++j) addInto(verts[f.v[j]].q, q);
    }
    for (size_t fi = 0; fi < faces.size(); ++fi) refreshCosts(fi);
  }
  void refreshCosts(size_t fi) {
    const Face& f = faces[fi];
    for (int j = 0; j < 3; ++j) {
      V3 unused;
      cost[3 * fi + j] = contraction(verts[f.v[j]], verts[f.v[(j + 1) % 3]], &unused);
    }
  }
  void dropDeletedFaces() {
    size_t keep = 0;
    for (size_t fi = 0; fi < faces.size(); ++fi) {
      if (flag[fi] & kDeleted) continue;
      if (keep != fi) {
        faces[keep] = faces[fi];
        for (int j = 0; j < 3; ++j) cost[3 * keep + j] = cost[3 * fi + j];
      }
      ++keep;
    }
    faces.resize(keep);
    cost.resize(3 * keep);
    flag.assign(keep, 0);
  }
  void rebuildIncidence() {
    for (Vertex& v : verts) v.faces.clear();
    for (size_t i = 0; i < faces.size(); ++i)
      for (int j = 0; j < 3; ++j) verts[faces[i].v[j]].faces.push_back((int)i);
  }
  std::vector<int> sharedFaces(int a, int b) const {
    std::vector<int> out;
    for (int fa : verts[a].faces)
      for (int fb : verts[b].faces)
        if (fa == fb) out.push_back(fa);
    return out;
  }
  // a vertex is on the boundary when an edge at it has a single face (identifySubBoundaries, whole range in one thread)
  void markBoundaries() {
    for (Vertex& v : verts) v.boundary = false;
    for (int i = 0; i < (int)verts.size(); ++i) {
      if (verts[i].boundary) continue;
      if (verts[i].faces.size() == 1) {
        verts[i].boundary = true;
        continue;
      }
      bool border = false;
      std::set<int> seen;
      for (int fi : verts[i].faces)
        for (int j = 0; j < 3; ++j) {
          const int o = faces[fi].v[j];
          if (o == i || !seen.insert(o).second) continue;
          if (verts[o].faces.size() == 1 || sharedFaces(i, o).size() == 1) {
            verts[o].boundary = true;
            border = true;
          }
        }
      if (border) verts[i].boundary = true;
    }
  }
  double costPercentile(float strictness) const {
    std::vector<double> all(cost);  // called right after the compaction: every face is live
    const int at = strictness * (all.size() - 1);  // float * size_t -> float -> int, as written in getThreshold
    std::nth_element(all.begin(), all.begin() + at, all.end());
    return all[at];
  }
  // would moving vertex a (edge a-b contracting) to p flip the normal of a face around a?
  bool flips(const V3& p, int a, int b) {
    for (size_t k = 0; k < verts[a].faces.size(); ++k) {
      if (flag[verts[a].faces[k]] & kDeleted) continue;
      const Face& f = faces[verts[a].faces[k]];
      int at = 0;
      for (int j = 0; j < 3; ++j)
        if (f.v[j] == a) {
          at = j;
          break;
        }
      const int i0 = f.v[(at + 1) % 3], i1 = f.v[(at + 2) % 3];
      if (i0 == b || i1 == b) continue;  // a face of the contracting edge itself
      const V3 e0 = unit(sub(verts[i0].p, p)), e1 = unit(sub(verts[i1].p, p));
      if (dot(unit(cross(e0, e1)), f.normal) < 0) return true;
    }
    return false;
  }
  void contract(int a, int b, const V3& p) {  // vertex a becomes the merged vertex
    verts[a].p = p;
    addInto(verts[a].q, verts[b].q);
    std::vector<int> around(verts[a].faces);
    around.insert(around.end(), verts[b].faces.begin(), verts[b].faces.end());
    for (int fi : around) {
      if (flag[fi] & kDeleted) continue;
      Face& f = faces[fi];
      for (int j = 0; j < 3; ++j)
        if (f.v[j] == a || f.v[j] == b) {
          f.v[j] = a;
          flag[fi] |= kTouched;
          break;
        }
      refreshCosts(fi);
    }
  }
  void compact() {
    std::vector<char> live(verts.size(), 0);
    dropDeletedFaces();
    for (const Face& f : faces)
      for (int j = 0; j < 3; ++j) live[f.v[j]] = 1;
    std::vector<int> renumber(verts.size(), -1);
    int next = 0;
    for (size_t i = 0; i < verts.size(); ++i)
      if (live[i]) {
        renumber[i] = next;
        verts[next++].p = verts[i].p;
      }
    verts.resize(next);
    for (Face& f : faces)
      for (int j = 0; j < 3; ++j) f.v[j] = renumber[f.v[j]];
  }

  // MeshSimplifier::simplify
  void run(int facesOut, float strictness, bool removeBoundaryEdges) {
    initialQuadrics();
    const int facesIn = (int)faces.size();
    int gone = 0, gonePrev = 0, stuck = 0, iteration = 0;
    double threshold = 0;
    // A sweep that contracts nothing leaves the mesh as it found it, and every edge it tried (all edges with cost <= its
    // threshold) failed for reasons that do not depend on the threshold (boundary rules, flipped normals).  The next sweep
    // then needs neither the compaction nor the incidence lists rebuilt, and only has to try the edges the larger
    // threshold newly admits; trying the others again is pure (no side effects) and would fail again.
    bool unchanged = false;  // the previous sweep contracted nothing
    double failedUpTo = 0;   // ... and every edge with cost <= failedUpTo was tried in it
    while ((int)faces.size() > facesOut) {
      if (!unchanged) {
        dropDeletedFaces();
        rebuildIncidence();
      }
      if (iteration == 0) markBoundaries();
      if (iteration == 0 || gonePrev != gone) {
        threshold = costPercentile(strictness);
        stuck = 0;
      } else {
        threshold *= 2 * ++stuck;  // nothing was contracted in the last sweep: open the threshold
        if (std::isinf(threshold)) break;
        // A threshold of exactly 0 (exactly planar patches: constant-disparity regions) or NaN never grows: the
        // reference's loop spins forever there (MeshSimplifier.cpp:505-512, reproduced with its own code in
        // tests/test_mesh.py).  Nothing more can be contracted under the reference's rules, so stop with the mesh as it is.
        if (!(threshold != 0) || threshold != threshold) break;
      }
      gonePrev = gone;
      bool skipKnown = unchanged;
      double maxCost = -std::numeric_limits<double>::infinity();
      for (size_t fi = 0; fi < faces.size(); ++fi) {
        // by index: contract() may not grow `faces`, but it writes through references into it
        if (flag[fi]) continue;  // deleted or touched in this sweep
        for (int j = 0; j < 3; ++j) {
          const double c = cost[3 * fi + j];
          if (!(c <= maxCost)) maxCost = c > maxCost ? c : std::numeric_limits<double>::quiet_NaN();
          if (c > threshold) continue;
          if (skipKnown && c <= failedUpTo) continue;
          const int a = faces[fi].v[j], b = faces[fi].v[(j + 1) % 3];
          if (verts[a].boundary != verts[b].boundary) continue;
          if (!removeBoundaryEdges && (verts[a].boundary || verts[b].boundary)) continue;
          V3 p;
          contraction(verts[a], verts[b], &p);
          if (flips(p, a, b) || flips(p, b, a)) continue;
          const std::vector<int> shared = sharedFaces(a, b);
          for (int s : shared) flag[s] |= kDeleted;
          gone += (int)shared.size();
          contract(a, b, p);
          skipKnown = false;  // the mesh moved: what failed before may succeed now
          break;
        }
        if (facesIn - gone <= facesOut) break;
      }
      ++iteration;
      unchanged = gone == gonePrev;
      if (unchanged) {
        failedUpTo = threshold;
        // every edge there is was tried and failed: larger thresholds change nothing until the reference's loop ends at inf
        if (maxCost <= threshold) break;
      }
    }
    compact();
  }
};

}  // namespace simplify
}  // namespace derp
