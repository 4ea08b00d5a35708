// Sparse 3-D reconstruction from monocular feature tracks, with the reference's interface
// (viso/reconstruction.h:40-67): feed it the flow matches of every frame pair and the egomotion, it links matches into
// tracks (via the feature indices i1p -> i1c), and when a track ends it triangulates the point, refines it by
// Gauss-Newton over all observations and keeps it if it passes the type / distance / ray-angle tests.
// Host code: a few hundred tracks per frame, microseconds of work next to the matcher.
// Own layout: frames carry absolute numbers instead of the reference's frames_ago counters and raw pointers; tracks
// live in a vector that is compacted in order, which preserves the reference's output order of points.
#ifndef VISOB_RECONSTRUCTION_H
#define VISOB_RECONSTRUCTION_H
#include <deque>
#include <vector>
#include "matcher.h"
#include "matrix.h"
#include "point3d.h"

Point3d affineTransform(const Matrix& M, const Point3d& p);      // rows 0..2 of M applied to (p, 1)  (matrix.cpp:37-44)

class Reconstruction {
public:
  Reconstruction();
  ~Reconstruction();

  // intrinsics (fu = fv = f); must be called once.  Also fixes the camera -> road transform used by the point types
  // (pitch -0.08 rad, height 1.6 m: reconstruction.cpp:37-48).
  void setCalibration(FLOAT f, FLOAT cu, FLOAT cv);

  // p_matched: flow matches of the newest frame pair; Tr: motion previous -> current camera coordinates.
  // point_type: 0 everything, 1 road and above, 2 only above the road; min_track_length in frames;
  // max_dist in metres from the camera; min_angle between the first and last viewing ray in degrees.
  void update(std::vector<Matcher::p_match> p_matched, Matrix Tr, int32_t point_type = 1, int32_t min_track_length = 2,
              double max_dist = 30, double min_angle = 2);

  // points of all finished tracks, in current camera coordinates
  const std::vector<Point3d>& getPoints() { return points; }

private:
  struct Obs { float u, v; };
  struct Frame {
    Matrix fwd, inv, proj;      // frame -> current camera, its inverse, K * inv[0:3, 0:4]
    int64_t id;                 // absolute frame number
    int32_t track_count;        // live tracks that started here
  };
  struct Track {
    std::vector<Obs> pixels;
    int64_t first_id;
    int32_t last_idx;
    bool refreshed;
  };
  static const unsigned max_track_length = 6;

  Frame& frame(int64_t id) { return frames[(size_t)(id - frames.front().id)]; }
  bool initPoint(const Track& t, Point3d& p);
  bool refinePoint(const Track& t, Point3d& p);
  int32_t pointType(const Track& t, const Point3d& p);
  double pointDistance(const Track& t, const Point3d& p);
  double rayAngle(const Track& t, const Point3d& p);

  Matrix K, Tr_cam_road;
  std::vector<Track> tracks;
  std::vector<Point3d> points;
  std::deque<Frame> frames;
  int64_t next_id;
};
#endif
