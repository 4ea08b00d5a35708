#include "viso.h"

#include <cmath>
#include <cstdlib>
#include <numeric>

VisualOdometry::VisualOdometry(parameters param) : sample_generator(71), param(param) {
  J = 0; p_observe = 0; p_predict = 0;
  matcher = new Matcher(param.match);
  owns_matcher = true;
  Tr_delta = Matrix::eye(4);
  Tr_valid = false;
  srand(0);                  // as the reference does (viso.cpp:35)
  matcher->seedShuffle(0);   // ... and the matcher's own copy of that generator, which bucketFeatures draws from
}

VisualOdometry::~VisualOdometry() { if (owns_matcher) delete matcher; }

void VisualOdometry::adoptMatcher(Matcher* external) {
  if (owns_matcher) delete matcher;
  matcher = external;
  owns_matcher = false;
  matcher->seedShuffle(0);     // same state the own matcher had (viso.cpp:35)
}

bool VisualOdometry::updateMotion() {
  std::vector<double> tr = estimateMotion(p_matched);
  if (tr.size() != 6) return false;
  Tr_delta = transformationVectorToMatrix(tr);
  Tr_valid = true;
  return true;
}

// rotation Rx(rx) * Ry(ry) * Rz(rz) with translation (viso.cpp:59-84)
Matrix VisualOdometry::transformationVectorToMatrix(std::vector<double> tr) {
  Matrix R = Matrix::rotMatX(tr[0]) * Matrix::rotMatY(tr[1]) * Matrix::rotMatZ(tr[2]);
  Matrix Tr = Matrix::eye(4);
  Tr.setMat(R, 0, 0);
  for (int i = 0; i < 3; i++) Tr.val[i][3] = tr[3 + i];
  return Tr;
}

std::vector<int> VisualOdometry::getRandomSample(unsigned N, unsigned num) {
  std::vector<int> pool(N);
  std::iota(pool.begin(), pool.end(), 0);
  for (unsigned i = 0; i < num; i++) {
    std::uniform_int_distribution<unsigned> pick(i, N - 1);
    std::swap(pool.at(i), pool.at(pick(sample_generator)));
  }
  pool.resize(num);
  return pool;
}
