// Stereo visual odometry with the reference's interface (viso/viso_stereo.h:27-84): quad matching with the
// motion-predicted search window (GPU), then a 3-point RANSAC over Gauss-Newton minimisations of the reprojection error
// and a final refinement on all inliers (host, double precision).
#ifndef VISOB_VISO_STEREO_H
#define VISOB_VISO_STEREO_H
#include "viso.h"

class VisualOdometryStereo : public VisualOdometry {
public:
  struct parameters : public VisualOdometry::parameters {
    double base;              // baseline (meters)
    int32_t ransac_iters;     // number of RANSAC iterations
    double inlier_threshold;  // reprojection inlier threshold (pixels)
    bool reweighting;         // lower border weights (more robust to calibration errors)
    parameters() { base = 1.0; ransac_iters = 200; inlier_threshold = 2.0; reweighting = true; }
  };

  VisualOdometryStereo(parameters param);
  ~VisualOdometryStereo();

  // returns false if an error occurred; valid after two calls
  bool process(uint8_t* I1, uint8_t* I2, uint32_t* dims, bool replace = false);
  using VisualOdometry::process;

private:
  enum result { UPDATED, FAILED, CONVERGED };
  std::vector<double> estimateMotion(std::vector<Matcher::p_match> p_matched);
  result updateParameters(const std::vector<Matcher::p_match>& p_matched, const std::vector<int32_t>& active, std::vector<double>& tr,
                          double step_size, double eps);
  void residualsAndJacobian(const std::vector<Matcher::p_match>& p_matched, const std::vector<int32_t>& active,
                            const std::vector<double>& tr, bool want_jacobian);
  std::vector<int32_t> getInlier(const std::vector<Matcher::p_match>& p_matched, const std::vector<double>& tr);

  std::vector<double> X, Y, Z;                     // 3-D points of the previous stereo pair
  std::vector<double> Jac, predict, observe, residual;
  parameters param;
};
#endif
