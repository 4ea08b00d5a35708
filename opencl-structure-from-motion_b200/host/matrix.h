// Small dense matrix type with the public interface of the reference's Matrix (viso/matrix.h:62-144): same member
// names (val, m, n), same operators and static helpers, so code written against libviso2 compiles unchanged.
// Host utility only -- nothing here is on the GPU path.  The implementation is independent of the reference's:
// contiguous storage behind the val[i][j] row pointers, and svd() is a one-sided Jacobi (Hestenes) SVD with the
// reference's output conventions (descending singular values, column signs flipped towards "mostly positive",
// U padded to m x m) instead of the Numerical-Recipes svdcmp of matrix.cpp:586-814.
#ifndef VISOB_MATRIX_H
#define VISOB_MATRIX_H
#include <stdint.h>
#include <iostream>
#include <vector>

typedef double FLOAT;

class Matrix {
public:
  Matrix();
  Matrix(const int32_t m, const int32_t n);
  Matrix(const int32_t m, const int32_t n, const FLOAT* val_);
  Matrix(const Matrix& M);
  ~Matrix();
  Matrix& operator=(const Matrix& M);

  void getData(FLOAT* val_, int32_t i1 = 0, int32_t j1 = 0, int32_t i2 = -1, int32_t j2 = -1);
  Matrix getMat(int32_t i1, int32_t j1, int32_t i2 = -1, int32_t j2 = -1);
  void setMat(const Matrix& M, const int32_t i, const int32_t j);
  void setVal(FLOAT s, int32_t i1 = 0, int32_t j1 = 0, int32_t i2 = -1, int32_t j2 = -1);
  void setDiag(FLOAT s, int32_t i1 = 0, int32_t i2 = -1);
  void zero();
  Matrix extractCols(std::vector<int> idx);

  static Matrix eye(const int32_t m);
  void eye();
  static Matrix diag(const Matrix& M);
  static Matrix reshape(const Matrix& M, int32_t m, int32_t n);
  static Matrix rotMatX(const FLOAT& angle);
  static Matrix rotMatY(const FLOAT& angle);
  static Matrix rotMatZ(const FLOAT& angle);

  Matrix operator+(const Matrix& M);
  Matrix operator-(const Matrix& M);
  Matrix operator*(const Matrix& M);
  Matrix operator*(const FLOAT& s);
  Matrix operator/(const Matrix& M);
  Matrix operator/(const FLOAT& s);
  Matrix operator-();
  Matrix operator~();
  FLOAT l2norm();
  FLOAT mean();

  static Matrix cross(const Matrix& a, const Matrix& b);
  static Matrix inv(const Matrix& M);
  bool inv();
  FLOAT det();
  bool solve(const Matrix& M, FLOAT eps = 1e-20);
  bool lu(int32_t* idx, FLOAT& d, FLOAT eps = 1e-20);
  void svd(Matrix& U, Matrix& W, Matrix& V);

  friend std::ostream& operator<<(std::ostream& out, const Matrix& M);

  FLOAT** val;
  int32_t m, n;

private:
  void allocateMemory(const int32_t m_, const int32_t n_);
  void releaseMemory();
};

#endif
