// See matrix.h.  Error convention follows the reference (message on cerr, exit(0); viso/matrix.cpp:100-105).
#include "matrix.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

using std::cerr;
using std::endl;

static void die(const char* what, int m1, int n1, int m2 = -1, int n2 = -1) {
  cerr << "ERROR: " << what << " (" << m1 << "x" << n1 << ")";
  if (m2 >= 0) cerr << " and (" << m2 << "x" << n2 << ")";
  cerr << endl;
  exit(0);
}

void Matrix::allocateMemory(const int32_t m_, const int32_t n_) {
  m = m_ < 0 ? 0 : m_;
  n = n_ < 0 ? 0 : n_;
  if (m == 0 || n == 0) { val = 0; return; }
  val = (FLOAT**)malloc(m * sizeof(FLOAT*));
  val[0] = (FLOAT*)calloc((size_t)m * n, sizeof(FLOAT));      // one slab, row pointers into it
  for (int32_t i = 1; i < m; i++) val[i] = val[i - 1] + n;
}
void Matrix::releaseMemory() {
  if (val) { free(val[0]); free(val); }
  val = 0;
}

Matrix::Matrix() : val(0), m(0), n(0) {}
Matrix::Matrix(const int32_t m_, const int32_t n_) { allocateMemory(m_, n_); }
Matrix::Matrix(const int32_t m_, const int32_t n_, const FLOAT* val_) {
  allocateMemory(m_, n_);
  if (val) memcpy(val[0], val_, (size_t)m * n * sizeof(FLOAT));
}
Matrix::Matrix(const Matrix& M) {
  allocateMemory(M.m, M.n);
  if (val) memcpy(val[0], M.val[0], (size_t)m * n * sizeof(FLOAT));
}
Matrix::~Matrix() { releaseMemory(); }
Matrix& Matrix::operator=(const Matrix& M) {
  if (this != &M) {
    if (M.m != m || M.n != n || (val == 0) != (M.val == 0)) { releaseMemory(); allocateMemory(M.m, M.n); }
    if (val) memcpy(val[0], M.val[0], (size_t)m * n * sizeof(FLOAT));
  }
  return *this;
}

void Matrix::getData(FLOAT* val_, int32_t i1, int32_t j1, int32_t i2, int32_t j2) {
  if (i2 == -1) i2 = m - 1;
  if (j2 == -1) j2 = n - 1;
  for (int32_t i = i1; i <= i2; i++)
    for (int32_t j = j1; j <= j2; j++) *val_++ = val[i][j];
}
Matrix Matrix::getMat(int32_t i1, int32_t j1, int32_t i2, int32_t j2) {
  if (i2 == -1) i2 = m - 1;
  if (j2 == -1) j2 = n - 1;
  if (i1 < 0 || i2 >= m || j1 < 0 || j2 >= n || i2 < i1 || j2 < j1) die("Cannot get that submatrix of a matrix of size", m, n);
  Matrix M(i2 - i1 + 1, j2 - j1 + 1);
  for (int32_t i = 0; i < M.m; i++) memcpy(M.val[i], val[i1 + i] + j1, M.n * sizeof(FLOAT));
  return M;
}
void Matrix::setMat(const Matrix& M, const int32_t i1, const int32_t j1) {
  if (i1 < 0 || j1 < 0 || i1 + M.m > m || j1 + M.n > n) die("Cannot set a submatrix of size", M.m, M.n, m, n);
  for (int32_t i = 0; i < M.m; i++) memcpy(val[i1 + i] + j1, M.val[i], M.n * sizeof(FLOAT));
}
void Matrix::setVal(FLOAT s, int32_t i1, int32_t j1, int32_t i2, int32_t j2) {
  if (i2 == -1) i2 = m - 1;
  if (j2 == -1) j2 = n - 1;
  if (i2 < i1 || j2 < j1) die("setVal indices must be ordered for a matrix of size", m, n);
  for (int32_t i = i1; i <= i2; i++) std::fill(val[i] + j1, val[i] + j2 + 1, s);
}
void Matrix::setDiag(FLOAT s, int32_t i1, int32_t i2) {
  if (i2 == -1) i2 = std::min(m - 1, n - 1);
  for (int32_t i = i1; i <= i2; i++) val[i][i] = s;
}
void Matrix::zero() { if (val) setVal(0); }
Matrix Matrix::extractCols(std::vector<int> idx) {
  Matrix M(m, (int32_t)idx.size());
  for (int32_t j = 0; j < M.n; j++)
    if (idx[j] < n)
      for (int32_t i = 0; i < m; i++) M.val[i][j] = val[i][idx[j]];
  return M;
}
Matrix Matrix::eye(const int32_t m) {
  Matrix M(m, m);
  M.setDiag(1);
  return M;
}
void Matrix::eye() {
  zero();
  setDiag(1);
}
Matrix Matrix::diag(const Matrix& M) {
  const bool colv = M.m > 1 && M.n == 1, rowv = M.m == 1 && M.n > 1;
  if (!colv && !rowv) die("Trying to create diagonal matrix from vector of size", M.m, M.n);
  const int32_t k = colv ? M.m : M.n;
  Matrix D(k, k);
  for (int32_t i = 0; i < k; i++) D.val[i][i] = colv ? M.val[i][0] : M.val[0][i];
  return D;
}
Matrix Matrix::reshape(const Matrix& M, int32_t m_, int32_t n_) {
  if (M.m * M.n != m_ * n_) die("Trying to reshape a matrix of size", M.m, M.n, m_, n_);
  Matrix R(m_, n_);
  if (R.val) memcpy(R.val[0], M.val[0], (size_t)m_ * n_ * sizeof(FLOAT));   // both row-major: element k -> element k
  return R;
}
static Matrix rot3(int a, int b, FLOAT angle) {    // rotation in the (a,b) plane of a 3x3 identity
  Matrix R = Matrix::eye(3);
  const FLOAT s = sin(angle), c = cos(angle);
  R.val[a][a] = c; R.val[a][b] = -s; R.val[b][a] = s; R.val[b][b] = c;
  return R;
}
Matrix Matrix::rotMatX(const FLOAT& angle) { return rot3(1, 2, angle); }
Matrix Matrix::rotMatY(const FLOAT& angle) { return rot3(2, 0, angle); }
Matrix Matrix::rotMatZ(const FLOAT& angle) { return rot3(0, 1, angle); }

Matrix Matrix::operator+(const Matrix& B) {
  if (m != B.m || n != B.n) die("Trying to add matrices of size", m, n, B.m, B.n);
  Matrix C(m, n);
  for (int32_t k = 0; k < m * n; k++) C.val[0][k] = val[0][k] + B.val[0][k];
  return C;
}
Matrix Matrix::operator-(const Matrix& B) {
  if (m != B.m || n != B.n) die("Trying to subtract matrices of size", m, n, B.m, B.n);
  Matrix C(m, n);
  for (int32_t k = 0; k < m * n; k++) C.val[0][k] = val[0][k] - B.val[0][k];
  return C;
}
Matrix Matrix::operator*(const Matrix& B) {
  if (n != B.m) die("Trying to multiply matrices of size", m, n, B.m, B.n);
  Matrix C(m, B.n);
  for (int32_t i = 0; i < m; i++)
    for (int32_t j = 0; j < B.n; j++) {
      FLOAT s = 0;
      for (int32_t k = 0; k < n; k++) s += val[i][k] * B.val[k][j];
      C.val[i][j] = s;
    }
  return C;
}
Matrix Matrix::operator*(const FLOAT& s) {
  Matrix C(m, n);
  for (int32_t k = 0; k < m * n; k++) C.val[0][k] = val[0][k] * s;
  return C;
}
Matrix Matrix::operator/(const Matrix& B) {
  // element-wise; a column / row vector divisor is broadcast; zero divisors leave 0 (viso/matrix.cpp:296-326)
  const bool same = m == B.m && n == B.n, colv = m == B.m && B.n == 1, rowv = n == B.n && B.m == 1;
  if (!same && !colv && !rowv) die("Trying to divide matrices of size", m, n, B.m, B.n);
  Matrix C(m, n);
  for (int32_t i = 0; i < m; i++)
    for (int32_t j = 0; j < n; j++) {
      const FLOAT d = same ? B.val[i][j] : (colv ? B.val[i][0] : B.val[0][j]);
      if (d != 0) C.val[i][j] = val[i][j] / d;
    }
  return C;
}
Matrix Matrix::operator/(const FLOAT& s) {
  if (fabs(s) < 1e-20) { cerr << "ERROR: Trying to divide by zero!" << endl; exit(0); }
  Matrix C(m, n);
  for (int32_t k = 0; k < m * n; k++) C.val[0][k] = val[0][k] / s;
  return C;
}
Matrix Matrix::operator-() {
  Matrix C(m, n);
  for (int32_t k = 0; k < m * n; k++) C.val[0][k] = -val[0][k];
  return C;
}
Matrix Matrix::operator~() {
  Matrix C(n, m);
  for (int32_t i = 0; i < m; i++)
    for (int32_t j = 0; j < n; j++) C.val[j][i] = val[i][j];
  return C;
}
FLOAT Matrix::l2norm() {
  FLOAT s = 0;
  for (int32_t k = 0; k < m * n; k++) s += val[0][k] * val[0][k];
  return sqrt(s);
}
FLOAT Matrix::mean() {
  FLOAT s = 0;
  for (int32_t k = 0; k < m * n; k++) s += val[0][k];
  return s / (FLOAT)(m * n);
}
Matrix Matrix::cross(const Matrix& a, const Matrix& b) {
  if (a.m != 3 || a.n != 1 || b.m != 3 || b.n != 1) die("Cross product vectors must be of size (3x1), got", a.m, a.n, b.m, b.n);
  Matrix c(3, 1);
  c.val[0][0] = a.val[1][0] * b.val[2][0] - a.val[2][0] * b.val[1][0];
  c.val[1][0] = a.val[2][0] * b.val[0][0] - a.val[0][0] * b.val[2][0];
  c.val[2][0] = a.val[0][0] * b.val[1][0] - a.val[1][0] * b.val[0][0];
  return c;
}

// LU decomposition with partial pivoting (Crout, implicit row scaling), in place.  idx receives the row
// permutation, d the permutation parity.  Returns false on a singular matrix.
bool Matrix::lu(int32_t* idx, FLOAT& d, FLOAT eps) {
  if (m != n) die("Trying to LU decompose a matrix of size", m, n);
  std::vector<FLOAT> scale(n);
  d = 1;
  for (int32_t i = 0; i < n; i++) {
    FLOAT big = 0;
    for (int32_t j = 0; j < n; j++) big = std::max(big, fabs(val[i][j]));
    if (big == 0) return false;
    scale[i] = 1.0 / big;
  }
  for (int32_t j = 0; j < n; j++) {
    for (int32_t i = 0; i < j; i++) {
      FLOAT s = val[i][j];
      for (int32_t k = 0; k < i; k++) s -= val[i][k] * val[k][j];
      val[i][j] = s;
    }
    FLOAT big = 0;
    int32_t imax = j;
    for (int32_t i = j; i < n; i++) {
      FLOAT s = val[i][j];
      for (int32_t k = 0; k < j; k++) s -= val[i][k] * val[k][j];
      val[i][j] = s;
      const FLOAT merit = scale[i] * fabs(s);
      if (merit >= big) { big = merit; imax = i; }
    }
    if (imax != j) {
      for (int32_t k = 0; k < n; k++) std::swap(val[imax][k], val[j][k]);
      d = -d;
      scale[imax] = scale[j];
    }
    idx[j] = imax;
    if (j != n - 1) {
      if (fabs(val[j][j]) < eps) return false;
      const FLOAT piv = 1.0 / val[j][j];
      for (int32_t i = j + 1; i < n; i++) val[i][j] *= piv;
    }
  }
  return true;
}
FLOAT Matrix::det() {
  if (m != n) die("Trying to compute determinant of a matrix of size", m, n);
  Matrix A(*this);
  std::vector<int32_t> idx(m);
  FLOAT d;
  if (!A.lu(idx.data(), d)) return 0;
  for (int32_t i = 0; i < m; i++) d *= A.val[i][i];
  return d;
}
// Gauss-Jordan elimination with full pivoting: on return *this holds the solution X of M*X = B (B = old *this)
// and M (cast away const, as in the reference) its inverse.
bool Matrix::solve(const Matrix& M, FLOAT eps) {
  Matrix& A = const_cast<Matrix&>(M);
  if (A.m != A.n || A.m != m || n < 1) die("Trying to solve a linear system with matrices of size", A.m, A.n, m, n);
  const int32_t N = A.m;
  std::vector<int32_t> rowi(N), coli(N), used(N, 0);
  for (int32_t step = 0; step < N; step++) {
    FLOAT big = 0;
    int32_t pr = 0, pc = 0;
    for (int32_t r = 0; r < N; r++) {
      if (used[r] == 1) continue;
      for (int32_t c = 0; c < N; c++)
        if (used[c] == 0 && fabs(A.val[r][c]) >= big) { big = fabs(A.val[r][c]); pr = r; pc = c; }
    }
    used[pc]++;
    if (pr != pc) {
      for (int32_t k = 0; k < N; k++) std::swap(A.val[pr][k], A.val[pc][k]);
      for (int32_t k = 0; k < n; k++) std::swap(val[pr][k], val[pc][k]);
    }
    rowi[step] = pr; coli[step] = pc;
    if (fabs(A.val[pc][pc]) < eps) return false;
    const FLOAT piv = 1.0 / A.val[pc][pc];
    A.val[pc][pc] = 1;
    for (int32_t k = 0; k < N; k++) A.val[pc][k] *= piv;
    for (int32_t k = 0; k < n; k++) val[pc][k] *= piv;
    for (int32_t r = 0; r < N; r++) {
      if (r == pc) continue;
      const FLOAT f = A.val[r][pc];
      A.val[r][pc] = 0;
      for (int32_t k = 0; k < N; k++) A.val[r][k] -= A.val[pc][k] * f;
      for (int32_t k = 0; k < n; k++) val[r][k] -= val[pc][k] * f;
    }
  }
  for (int32_t step = N - 1; step >= 0; step--)
    if (rowi[step] != coli[step])
      for (int32_t k = 0; k < N; k++) std::swap(A.val[k][rowi[step]], A.val[k][coli[step]]);
  return true;
}
bool Matrix::inv() {
  if (m != n) die("Trying to invert a matrix of size", m, n);
  Matrix A(*this);
  eye();
  return solve(A);
}
Matrix Matrix::inv(const Matrix& M) {
  if (M.m != M.n) die("Trying to invert a matrix of size", M.m, M.n);
  Matrix A(M), B = eye(M.m);
  B.solve(A);
  return B;
}

// One-sided Jacobi SVD.  G = A is rotated column by column until all columns are mutually orthogonal; then
// sigma_j = |g_j|, u_j = g_j / sigma_j, and the accumulated rotations are V.  Output conventions as in the
// reference (matrix.cpp:766-807): n singular values sorted in decreasing order (W keeps the first min(m,n)),
// each (u_j, v_j) pair negated when more than half of its entries are negative, U returned as m x m.
void Matrix::svd(Matrix& U2, Matrix& W, Matrix& V) {
  const int32_t M = m, N = n;
  std::vector<FLOAT> g((size_t)M * N), v((size_t)N * N, 0.0);     // column-major
  for (int32_t i = 0; i < M; i++)
    for (int32_t j = 0; j < N; j++) g[(size_t)j * M + i] = val[i][j];
  for (int32_t j = 0; j < N; j++) v[(size_t)j * N + j] = 1.0;
  const FLOAT tol = 1e-15;
  FLOAT total = 0;
  for (size_t k = 0; k < g.size(); k++) total += g[k] * g[k];
  const FLOAT tiny = 1e-28 * total;       // numerically null columns are left alone (rounding noise cannot be rotated away)
  bool converged = false;
  for (int sweep = 0; sweep < 60 && !converged; sweep++) {
    converged = true;
    for (int32_t p = 0; p < N - 1; p++)
      for (int32_t q = p + 1; q < N; q++) {
        FLOAT* gp = &g[(size_t)p * M];
        FLOAT* gq = &g[(size_t)q * M];
        FLOAT alpha = 0, beta = 0, gamma = 0;
        for (int32_t i = 0; i < M; i++) { alpha += gp[i] * gp[i]; beta += gq[i] * gq[i]; gamma += gp[i] * gq[i]; }
        if (alpha <= tiny || beta <= tiny || fabs(gamma) <= tol * sqrt(alpha * beta)) continue;
        converged = false;
        const FLOAT zeta = (beta - alpha) / (2.0 * gamma);
        const FLOAT t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const FLOAT c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int32_t i = 0; i < M; i++) { const FLOAT a = gp[i], b = gq[i]; gp[i] = c * a - s * b; gq[i] = s * a + c * b; }
        FLOAT* vp = &v[(size_t)p * N];
        FLOAT* vq = &v[(size_t)q * N];
        for (int32_t i = 0; i < N; i++) { const FLOAT a = vp[i], b = vq[i]; vp[i] = c * a - s * b; vq[i] = s * a + c * b; }
      }
  }
  if (!converged) cerr << "ERROR in SVD: No convergence in 60 Jacobi sweeps" << endl;   // cf. matrix.cpp:713-714
  std::vector<FLOAT> w(N);
  for (int32_t j = 0; j < N; j++) {
    FLOAT s = 0;
    for (int32_t i = 0; i < M; i++) s += g[(size_t)j * M + i] * g[(size_t)j * M + i];
    w[j] = sqrt(s);
  }
  std::vector<int32_t> order(N);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return w[a] > w[b]; });
  Matrix Un(M, N);
  V = Matrix(N, N);
  std::vector<FLOAT> ws(N);
  // left singular vectors: g_j / sigma_j; for a (numerically) zero singular value that quotient is noise, so the column
  // is completed to an orthonormal set instead (the pose recovery multiplies with the full U of a rank-2 matrix)
  const FLOAT wmax = N > 0 ? w[order[0]] : 0;
  for (int32_t k = 0; k < N && k < M; k++) {
    const int32_t j = order[k];
    if (w[j] > 1e-13 * wmax && w[j] > 0) {
      for (int32_t i = 0; i < M; i++) Un.val[i][k] = g[(size_t)j * M + i] / w[j];
      continue;
    }
    FLOAT best = -1;
    std::vector<FLOAT> cand(M), pick(M, 0.0);
    for (int32_t e = 0; e < M; e++) {
      for (int32_t i = 0; i < M; i++) cand[i] = i == e ? 1.0 : 0.0;
      for (int pass = 0; pass < 2; pass++)
        for (int32_t c = 0; c < k; c++) {
          FLOAT dot = 0;
          for (int32_t i = 0; i < M; i++) dot += Un.val[i][c] * cand[i];
          for (int32_t i = 0; i < M; i++) cand[i] -= dot * Un.val[i][c];
        }
      FLOAT nrm = 0;
      for (int32_t i = 0; i < M; i++) nrm += cand[i] * cand[i];
      if (nrm > best) { best = nrm; pick = cand; }
    }
    const FLOAT inv = best > 0 ? 1.0 / sqrt(best) : 0.0;
    for (int32_t i = 0; i < M; i++) Un.val[i][k] = pick[i] * inv;
  }
  for (int32_t k = 0; k < N; k++) {
    const int32_t j = order[k];
    ws[k] = w[j];
    int32_t neg = 0;
    for (int32_t i = 0; i < M; i++) {
      if (Un.val[i][k] < 0) neg++;
    }
    for (int32_t i = 0; i < N; i++) {
      V.val[i][k] = v[(size_t)j * N + i];
      if (V.val[i][k] < 0) neg++;
    }
    if (neg > (M + N) / 2) {
      for (int32_t i = 0; i < M; i++) Un.val[i][k] = -Un.val[i][k];
      for (int32_t i = 0; i < N; i++) V.val[i][k] = -V.val[i][k];
    }
  }
  W = Matrix(std::min(M, N), 1, ws.data());
  U2 = Matrix(M, M);
  U2.setMat(Un.getMat(0, 0, M - 1, std::min(M - 1, N - 1)), 0, 0);
}

std::ostream& operator<<(std::ostream& out, const Matrix& M) {
  if (M.m == 0 || M.n == 0) {
    out << "[empty matrix]";
  } else {
    char buffer[1024];
    for (int32_t i = 0; i < M.m; i++) {
      for (int32_t j = 0; j < M.n; j++) {
        snprintf(buffer, sizeof buffer, "%12.7f ", M.val[i][j]);
        out << buffer;
      }
      if (i < M.m - 1) out << endl;
    }
  }
  return out;
}
