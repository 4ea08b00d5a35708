// Reconstruction (reference behaviour: viso/reconstruction.cpp:27-343), see reconstruction.h.
#include "reconstruction.h"
#include <math.h>

using std::vector;

Point3d affineTransform(const Matrix& M, const Point3d& p) {
  // double accumulation on float coordinates, rounded to float on the way out (matrix.cpp:37-44)
  return Point3d((float)(p.x * M.val[0][0] + p.y * M.val[0][1] + p.z * M.val[0][2] + M.val[0][3]),
                 (float)(p.x * M.val[1][0] + p.y * M.val[1][1] + p.z * M.val[1][2] + M.val[1][3]),
                 (float)(p.x * M.val[2][0] + p.y * M.val[2][1] + p.z * M.val[2][2] + M.val[2][3]));
}

Reconstruction::Reconstruction() : next_id(0) { K = Matrix::eye(3); }
Reconstruction::~Reconstruction() {}

void Reconstruction::setCalibration(FLOAT f, FLOAT cu, FLOAT cv) {
  const FLOAT kd[9] = {f, 0, cu, 0, f, cv, 0, 0, 1};
  K = Matrix(3, 3, kd);
  const FLOAT pitch = -0.08, height = 1.6;
  Tr_cam_road = Matrix(4, 4);          // note: row 3 stays zero, only rows 0..2 are ever applied
  Tr_cam_road.val[0][0] = 1;
  Tr_cam_road.val[1][1] = +cos(pitch); Tr_cam_road.val[1][2] = -sin(pitch);
  Tr_cam_road.val[2][1] = +sin(pitch); Tr_cam_road.val[2][2] = +cos(pitch);
  Tr_cam_road.val[1][3] = -height;
}

void Reconstruction::update(vector<Matcher::p_match> p_matched, Matrix Tr, int32_t point_type, int32_t min_track_length,
                            double max_dist, double min_angle) {
  // everything already reconstructed moves into the new camera frame
  for (Point3d& p : points) p = affineTransform(Tr, p);
  // frames no live track starts in are dropped from the old end; the rest are re-expressed relative to the new frame
  while (!frames.empty() && frames.front().track_count == 0) frames.pop_front();
  for (Frame& f : frames) {
    f.fwd = Tr * f.fwd;
    f.inv = Matrix::inv(f.fwd);
    f.proj = K * f.inv.getMat(0, 0, 2, 3);
  }
  {
    Frame f;
    f.fwd = Matrix::eye(4); f.inv = Matrix::eye(4); f.proj = K * f.fwd.getMat(0, 0, 2, 3);
    f.id = next_id++; f.track_count = 0;
    frames.push_back(f);
  }
  const int64_t now = frames.back().id;

  // feature index of the previous frame -> track that ended on it (a later track wins, as in the reference's map)
  int32_t idx_max = 0;
  for (const Matcher::p_match& m : p_matched) if (m.i1p > idx_max) idx_max = m.i1p;
  for (const Track& t : tracks) if (t.last_idx > idx_max) idx_max = t.last_idx;
  vector<int32_t> by_index((size_t)idx_max + 1, -1);
  for (size_t k = 0; k < tracks.size(); k++) by_index[tracks[k].last_idx] = (int32_t)k;

  for (const Matcher::p_match& m : p_matched) {
    int32_t k = by_index[m.i1p];
    if (k < 0 || tracks[k].refreshed) {
      // new track: it starts with the observation in the previous image, but is anchored at the newest frame
      // (reconstruction.cpp:88-94) -- the projections used later are shifted by one frame accordingly
      Track t;
      t.first_id = now; t.last_idx = 0; t.refreshed = false;
      t.pixels.push_back(Obs{m.u1p, m.v1p});
      frames.back().track_count += 1;
      tracks.push_back(t);
      k = (int32_t)tracks.size() - 1;
    }
    Track& t = tracks[k];
    if (t.pixels.size() < max_track_length) {
      t.pixels.push_back(Obs{m.u1c, m.v1c});
      t.last_idx = m.i1c;
      t.refreshed = true;
    }
  }

  // tracks that were not continued end here: reconstruct, then drop them (order preserved)
  size_t keep = 0;
  for (size_t k = 0; k < tracks.size(); k++) {
    Track& t = tracks[k];
    if (t.refreshed) {
      t.refreshed = false;
      if (keep != k) tracks[keep] = std::move(t);
      keep++;
      continue;
    }
    if ((int32_t)t.pixels.size() >= min_track_length) {
      Point3d p;
      if (initPoint(t, p) && pointType(t, p) >= point_type && refinePoint(t, p) &&
          pointDistance(t, p) < max_dist && rayAngle(t, p) > min_angle)
        points.push_back(p);
    }
    frame(t.first_id).track_count -= 1;
  }
  tracks.resize(keep);
}

// linear triangulation from the first and the last observation: null vector of the 4x4 system
bool Reconstruction::initPoint(const Track& t, Point3d& p) {
  const Matrix& P1 = frame(t.first_id).proj;
  const Matrix& P2 = frames.back().proj;
  const Obs a = t.pixels.front(), b = t.pixels.back();
  Matrix J(4, 4), U, S, V;
  for (int32_t j = 0; j < 4; j++) {
    J.val[0][j] = P1.val[2][j] * a.u - P1.val[0][j];
    J.val[1][j] = P1.val[2][j] * a.v - P1.val[1][j];
    J.val[2][j] = P2.val[2][j] * b.u - P2.val[0][j];
    J.val[3][j] = P2.val[2][j] * b.v - P2.val[1][j];
  }
  J.svd(U, S, V);
  const float w = (float)V.val[3][3];
  if (fabs(w) < 1e-10) return false;            // point at infinity
  p = Point3d((float)(V.val[0][3] / w), (float)(V.val[1][3] / w), (float)(V.val[2][3] / w));
  return true;
}

// Gauss-Newton on the reprojection error over all observations of the track; observation k is predicted with the
// projection of frame first + k.  At most 22 steps; only a converged point (all updates below 1e-5) is accepted.
bool Reconstruction::refinePoint(const Track& t, Point3d& p) {
  const int32_t nf = (int32_t)t.pixels.size();
  vector<FLOAT> J(6 * (size_t)nf), res(2 * (size_t)nf);
  const size_t f0 = (size_t)(t.first_id - frames.front().id);
  for (int32_t iter = 0;; iter++) {
    for (int32_t k = 0; k < nf; k++) {
      const Matrix& P = frames[f0 + k].proj;
      const FLOAT a = P.val[0][0] * p.x + P.val[0][1] * p.y + P.val[0][2] * p.z + P.val[0][3];
      const FLOAT b = P.val[1][0] * p.x + P.val[1][1] * p.y + P.val[1][2] * p.z + P.val[1][3];
      const FLOAT c = P.val[2][0] * p.x + P.val[2][1] * p.y + P.val[2][2] * p.z + P.val[2][3];
      const FLOAT cc = c * c;
      if (cc < 1e-10) return false;
      for (int32_t j = 0; j < 3; j++) {
        J[6 * k + j] = (P.val[0][j] * c - P.val[2][j] * a) / cc;
        J[6 * k + 3 + j] = (P.val[1][j] * c - P.val[2][j] * b) / cc;
      }
      res[2 * k] = t.pixels[k].u - a / c;
      res[2 * k + 1] = t.pixels[k].v - b / c;
    }
    Matrix A(3, 3), B(3, 1);
    for (int32_t m = 0; m < 3; m++) {
      for (int32_t n = 0; n < 3; n++) {
        FLOAT s = 0;
        for (int32_t i = 0; i < 2 * nf; i++) s += J[3 * i + m] * J[3 * i + n];
        A.val[m][n] = s;
      }
      FLOAT s = 0;
      for (int32_t i = 0; i < 2 * nf; i++) s += J[3 * i + m] * res[i];
      B.val[m][0] = s;
    }
    if (!B.solve(A)) return false;
    p.x += B.val[0][0]; p.y += B.val[1][0]; p.z += B.val[2][0];
    if (fabs(B.val[0][0]) < 1e-5 && fabs(B.val[1][0]) < 1e-5 && fabs(B.val[2][0]) < 1e-5) return true;
    if (iter > 20) return false;
  }
}

// -1 not visible (closer than 1 m to either end camera), 0 below the road, 1 road, 2 obstacle
int32_t Reconstruction::pointType(const Track& t, const Point3d& p) {
  const Point3d x1c = affineTransform(frame(t.first_id).inv, p);
  const Point3d x2c = affineTransform(frames.back().inv, p);
  const Point3d x2r = affineTransform(Tr_cam_road, x2c);
  if (x1c.z <= 1 || x2c.z <= 1) return -1;
  if (x2r.y > 0.5) return 0;
  if (x2r.y > -1) return 1;
  return 2;
}

// distance to the camera position half way along the track
double Reconstruction::pointDistance(const Track& t, const Point3d& p) {
  const unsigned first_ago = (unsigned)(frames.back().id - t.first_id);
  const unsigned mid_ago = (first_ago + 0u + 1u) / 2u;
  const Matrix& fwd = frames[frames.size() - 1 - mid_ago].fwd;
  const double dx = fwd.val[0][3] - p.x, dy = fwd.val[1][3] - p.y, dz = fwd.val[2][3] - p.z;
  return sqrt(dx * dx + dy * dy + dz * dz);
}

// angle (degrees) between the viewing rays from the first and the last camera centre
double Reconstruction::rayAngle(const Track& t, const Point3d& p) {
  const Matrix& F1 = frame(t.first_id).fwd;
  const Matrix& F2 = frames.back().fwd;
  FLOAT v1[3], v2[3];
  const FLOAT pt[3] = {p.x, p.y, p.z};
  for (int i = 0; i < 3; i++) { v1[i] = F1.val[i][3] - pt[i]; v2[i] = F2.val[i][3] - pt[i]; }
  const FLOAT n1 = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
  const FLOAT n2 = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
  if (n1 < 1e-10 || n2 < 1e-10) return 1000;
  FLOAT dot = 0;
  for (int i = 0; i < 3; i++) dot += (v1[i] / n1) * (v2[i] / n2);
  return acos(fabs(dot)) * 180.0 / M_PI;
}
