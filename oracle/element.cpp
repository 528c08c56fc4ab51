// TEST INFRASTRUCTURE ONLY (see fe_oracle.hpp).  Reference element + per-element operators.
#include "fe_oracle.hpp"

#include <algorithm>
#include <stdexcept>

namespace feo {

// common/scale_polynomial.F90 Polynomial_GenLegendrePoly_sub: three-term recurrence
static void legendre(int nord, double x, double* P) {
  P[0] = 1.0;
  if (nord == 0) return;
  P[1] = x;
  for (int n = 2; n <= nord; ++n) P[n] = ((2 * n - 1) * x * P[n - 1] - (n - 1) * P[n - 2]) / n;
}

// Dense inverse by Gauss-Jordan with partial pivoting (the reference calls LAPACK dgetrf/dgetri,
// common/scale_linalgebra.F90:90-96).
static vec inverse(const vec& A, int n) {
  vec a(A), inv(size_t(n) * n, 0.0);
  for (int i = 0; i < n; ++i) inv[i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[r * n + c]) > std::fabs(a[piv * n + c])) piv = r;
    if (a[piv * n + c] == 0.0) throw std::runtime_error("singular matrix");
    if (piv != c)
      for (int k = 0; k < n; ++k) { std::swap(a[c * n + k], a[piv * n + k]); std::swap(inv[c * n + k], inv[piv * n + k]); }
    double d = 1.0 / a[c * n + c];
    for (int k = 0; k < n; ++k) { a[c * n + k] *= d; inv[c * n + k] *= d; }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      double f = a[r * n + c];
      if (f == 0.0) continue;
      for (int k = 0; k < n; ++k) { a[r * n + k] -= f * a[c * n + k]; inv[r * n + k] -= f * inv[c * n + k]; }
    }
  }
  return inv;
}

static vec matmul(const vec& A, const vec& B, int m, int k, int n) {
  vec C(size_t(m) * n, 0.0);
  for (int i = 0; i < m; ++i)
    for (int l = 0; l < k; ++l) {
      double a = A[i * k + l];
      for (int j = 0; j < n; ++j) C[i * n + j] += a * B[l * n + j];
    }
  return C;
}

// LGL nodes: the reference takes eigenvalues of the Jacobi matrix with LAPACK dstev
// (scale_polynomial.F90:330-372); here the same nodes are found as the roots of P'_N by Newton
// iteration from Chebyshev-Gauss-Lobatto guesses.  Both give the nodes to round-off.
static vec lgl_nodes(int N) {
  vec x(N + 1);
  x[0] = -1.0; x[N] = 1.0;
  vec P(N + 2);
  for (int i = 1; i < N; ++i) {
    double xi = -std::cos(M_PI * i / N);
    for (int it = 0; it < 100; ++it) {
      legendre(N, xi, P.data());
      double dP = N * (P[N - 1] - xi * P[N]) / (1.0 - xi * xi);           // P'_N
      double d2P = (2.0 * xi * dP - N * (N + 1.0) * P[N]) / (1.0 - xi * xi);  // P''_N
      double dx = dP / d2P;
      xi -= dx;
      if (std::fabs(dx) < 1e-16) break;
    }
    x[i] = xi;
  }
  // symmetrise (the eigen-solver result is symmetric to round-off as well)
  for (int i = 0; i <= N / 2; ++i) { double a = 0.5 * (x[N - i] - x[i]); x[i] = -a; x[N - i] = a; }
  if (N % 2 == 0) x[N / 2] = 0.0;
  return x;
}

void Element::init(int order, bool lumped_mass) {
  p = order; np = p + 1; Np = np * np * np; Nfp = np * np; NfpTot = 6 * Nfp; lumped = lumped_mass;
  x1d = lgl_nodes(p);
  // weights: scale_polynomial.F90 Polynomial_GenGaussLobattoPtIntWeight
  w1d.resize(np);
  vec Pn(size_t(np) * np);
  for (int i = 0; i < np; ++i) legendre(p, x1d[i], &Pn[i * np]);
  for (int i = 0; i < np; ++i) w1d[i] = 2.0 / (double(p) * (p + 1) * Pn[i * np + p] * Pn[i * np + p]);
  // V (scale_element_line.F90: V(n,l) = P_l(x_n) sqrt(l + 1/2))
  V1.resize(size_t(np) * np);
  for (int i = 0; i < np; ++i)
    for (int l = 0; l < np; ++l) V1[i * np + l] = Pn[i * np + l] * std::sqrt(l + 0.5);
  invV1 = inverse(V1, np);
  // D1D: scale_polynomial.F90 Polynomial_GenDLagrangePoly_lglpt; Dx1(n,l) = lr(l,n)
  D1D.assign(size_t(np) * np, 0.0);
  for (int n = 0; n < np; ++n) {
    double s = 0.0;
    for (int k = 0; k < np; ++k) {
      double v;
      if (k == 0 && n == 0) v = -0.25 * p * (p + 1);
      else if (k == p && n == p) v = 0.25 * p * (p + 1);
      else if (k == n) v = 0.0;
      else v = Pn[n * np + p] / (Pn[k * np + p] * (x1d[n] - x1d[k]));
      if (k != n) { s += v; D1D[n * np + k] = v; }
    }
    D1D[n * np + n] = -s;
  }
  // mass matrices (scale_element_base.F90 ElementBase_construct_MassMat: invM = V V^T)
  if (lumped) {
    invM1.assign(size_t(np) * np, 0.0); M1 = invM1;
    for (int i = 0; i < np; ++i) { M1[i * np + i] = w1d[i]; invM1[i * np + i] = 1.0 / w1d[i]; }
  } else {
    vec Vt(size_t(np) * np);
    for (int i = 0; i < np; ++i) for (int l = 0; l < np; ++l) Vt[l * np + i] = V1[i * np + l];
    invM1 = matmul(V1, Vt, np, np, np);
    M1 = inverse(invM1, np);
  }
  lift1d.resize(size_t(np) * 2);
  for (int m = 0; m < np; ++m) { lift1d[m * 2] = invM1[m * np + 0]; lift1d[m * 2 + 1] = invM1[m * np + p]; }
  // IntrpMat_VPOrdM1 (tensorprod3D.F90.erb:556-559)
  vec iv(invV1);
  for (int l = 0; l < np; ++l) iv[p * np + l] = 0.0;
  VPOrdM1 = matmul(V1, iv, np, np, np);
  filt_h.assign(size_t(np) * np, 0.0); for (int i = 0; i < np; ++i) filt_h[i * np + i] = 1.0;
  filt_v = filt_h;
  IntWeight.resize(Np);
  Fmask.resize(size_t(6) * Nfp);
  for (int k = 0; k < np; ++k) for (int j = 0; j < np; ++j) for (int i = 0; i < np; ++i)
    IntWeight[i + j * np + k * np * np] = w1d[i] * w1d[j] * w1d[k];
  auto nid = [&](int i, int j, int k) { return i + j * np + k * np * np; };
  for (int b = 0; b < np; ++b) for (int a = 0; a < np; ++a) {
    int fp = a + b * np;
    Fmask[0 * Nfp + fp] = nid(a, 0, b);
    Fmask[1 * Nfp + fp] = nid(p, a, b);
    Fmask[2 * Nfp + fp] = nid(a, p, b);
    Fmask[3 * Nfp + fp] = nid(0, a, b);
    Fmask[4 * Nfp + fp] = nid(a, b, 0);
    Fmask[5 * Nfp + fp] = nid(a, b, p);
  }
}

// element/scale_element_modalfilter.F90:204-236 (get_exp_filter) and ModalFilter_Init_line
static vec filter1d(const Element& e, double etac, double alpha, int ord) {
  int np = e.np, p = e.p;
  vec F(size_t(np) * np, 0.0);
  for (int m = 0; m < np; ++m) {
    double eta = double(m) / double(p), f = 1.0;
    if (eta > etac && m != 0) f = std::exp(-alpha * std::pow((eta - etac) / (1.0 - etac), ord));
    F[m * np + m] = f;
  }
  return matmul(e.V1, matmul(F, e.invV1, np, np, np), np, np, np);
}

void Element::setup_filter(double etac_h, double alpha_h, int ord_h, double etac_v, double alpha_v, int ord_v) {
  filt_h = filter1d(*this, etac_h, alpha_h, ord_h);
  filt_v = filter1d(*this, etac_v, alpha_v, ord_v);
}

vec Element::dmat_dense(int dir) const {
  vec D(size_t(Np) * Np, 0.0);
  for (int k = 0; k < np; ++k) for (int j = 0; j < np; ++j) for (int i = 0; i < np; ++i) {
    int n = i + j * np + k * np * np;
    for (int l = 0; l < np; ++l) {
      int m = dir == 0 ? l + j * np + k * np * np : dir == 1 ? i + l * np + k * np * np : i + j * np + l * np * np;
      int a = dir == 0 ? i : dir == 1 ? j : k;
      D[size_t(n) * Np + m] = D1D[a * np + l];
    }
  }
  return D;
}

// hexahedral.F90:331-400: Lift = invM * Emat with Emat(Fmask(:,f), face cols) = face mass matrix.
// invM = invM1 (x) invM1 (x) invM1 and the face mass matrix is M1 (x) M1, both built from their
// 1D factors here (the Kronecker identities hold exactly for the reference's V = V1 (x) V1 (x) V1).
vec Element::lift_dense() const {
  vec L(size_t(Np) * NfpTot, 0.0);
  for (int f = 0; f < 6; ++f)
    for (int b2 = 0; b2 < np; ++b2) for (int a2 = 0; a2 < np; ++a2) {       // face node (a2,b2) = column
      int col = f * Nfp + a2 + b2 * np;
      for (int k = 0; k < np; ++k) for (int j = 0; j < np; ++j) for (int i = 0; i < np; ++i) {
        int n = i + j * np + k * np * np;
        // (invM Emat)(n,col) = sum over face nodes (a1,b1) invM(n, node(a1,b1)) * Mface((a1,b1),(a2,b2))
        double s = 0.0;
        for (int b1 = 0; b1 < np; ++b1) for (int a1 = 0; a1 < np; ++a1) {
          int fi, fj, fk;
          switch (f) {
            case 0: fi = a1; fj = 0; fk = b1; break;
            case 1: fi = p; fj = a1; fk = b1; break;
            case 2: fi = a1; fj = p; fk = b1; break;
            case 3: fi = 0; fj = a1; fk = b1; break;
            case 4: fi = a1; fj = b1; fk = 0; break;
            default: fi = a1; fj = b1; fk = p; break;
          }
          double invm = invM1[i * np + fi] * invM1[j * np + fj] * invM1[k * np + fk];
          double mf = M1[a1 * np + a2] * M1[b1 * np + b2];
          s += invm * mf;
        }
        L[size_t(n) * NfpTot + col] = s;
      }
    }
  return L;
}

// ---- per-element kernels (tensorprod3D_kernel.F90.erb:55-159): left-to-right sums over the 1D index
void op_dx(const Element& e, const double* in, double* out) {
  const int np = e.np;
  for (int jk = 0; jk < np * np; ++jk)
    for (int i = 0; i < np; ++i) {
      double s = e.D1D[i * np] * in[jk * np];
      for (int l = 1; l < np; ++l) s += e.D1D[i * np + l] * in[l + jk * np];
      out[i + jk * np] = s;
    }
}
void op_dy(const Element& e, const double* in, double* out) {
  const int np = e.np;
  for (int k = 0; k < np; ++k) for (int j = 0; j < np; ++j) for (int i = 0; i < np; ++i) {
    double s = in[i + k * np * np] * e.D1D[j * np];
    for (int l = 1; l < np; ++l) s += in[i + l * np + k * np * np] * e.D1D[j * np + l];
    out[i + j * np + k * np * np] = s;
  }
}
void op_matz(const Element& e, const double* Mat, const double* in, double* out) {
  const int np = e.np;
  for (int k = 0; k < np; ++k) for (int j = 0; j < np; ++j) for (int i = 0; i < np; ++i) {
    double s = in[i + j * np] * Mat[k * np];
    for (int l = 1; l < np; ++l) s += in[i + j * np + l * np * np] * Mat[k * np + l];
    out[i + j * np + k * np * np] = s;
  }
}
void op_dz(const Element& e, const double* in, double* out) { op_matz(e, e.D1D.data(), in, out); }

void op_lift(const Element& e, const double* f, double* out) {
  const int np = e.np, Nfp = e.Nfp;
  const double* lw = e.lift1d.data();
  for (int k = 0; k < np; ++k) for (int j = 0; j < np; ++j) for (int i = 0; i < np; ++i)
    out[i + j * np + k * np * np] =
        lw[j * 2] * f[i + k * np] + lw[i * 2 + 1] * f[Nfp + j + k * np] + lw[j * 2 + 1] * f[2 * Nfp + i + k * np] +
        lw[i * 2] * f[3 * Nfp + j + k * np] + lw[k * 2] * f[4 * Nfp + i + j * np] + lw[k * 2 + 1] * f[5 * Nfp + i + j * np];
}

void op_div(const Element& e, const double* flux3, const double* del_flux, double* d4) {
  const int Np = e.Np;
  op_lift(e, del_flux, d4 + 3 * Np);
  op_dx(e, flux3, d4);
  op_dy(e, flux3 + Np, d4 + Np);
  op_dz(e, flux3 + 2 * Np, d4 + 2 * Np);
}

// tensorprod3D_kernel.F90.erb:468-582: x pass, y pass, z pass
void op_modal_filter(const Element& e, const double* in, double* work, double* out) {
  const int np = e.np;
  for (int jk = 0; jk < np * np; ++jk)
    for (int i = 0; i < np; ++i) {
      double s = e.filt_h[i * np] * in[jk * np];
      for (int l = 1; l < np; ++l) s += e.filt_h[i * np + l] * in[l + jk * np];
      out[i + jk * np] = s;
    }
  for (int k = 0; k < np; ++k) for (int j = 0; j < np; ++j) for (int i = 0; i < np; ++i) {
    double s = out[i + k * np * np] * e.filt_h[j * np];
    for (int l = 1; l < np; ++l) s += out[i + l * np + k * np * np] * e.filt_h[j * np + l];
    work[i + j * np + k * np * np] = s;
  }
  op_matz(e, e.filt_v.data(), work, out);
}

// ---------------------------------------------------------------- sparse matrix
// common/scale_sparsemat.F90:100-250 (Init: drop |a| <= eps), :439-474 / :554-634 (matmul)
void SparseMat::init(const double* A, int M_, int N_, double eps, bool ell_format) {
  M = M_; N = N_; ell = ell_format; nnz = 0; col_size = 0;
  std::vector<std::vector<std::pair<int, double>>> rows(M);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j)
      if (std::fabs(A[size_t(i) * N + j]) > eps) { rows[i].push_back({j, A[size_t(i) * N + j]}); ++nnz; }
  for (auto& r : rows) col_size = std::max<int>(col_size, int(r.size()));
  if (!ell) {
    rowPtr.assign(M + 1, 0);
    for (int i = 0; i < M; ++i) {
      rowPtr[i + 1] = rowPtr[i] + int(rows[i].size());
      for (auto& pr : rows[i]) { colIdx.push_back(pr.first); val.push_back(pr.second); }
    }
  } else {
    // ELL: slot-major storage l = i + k*M (scale_sparsemat.F90:172); padded with value 0 / column of row i
    val.assign(size_t(M) * col_size, 0.0);
    colIdx.assign(size_t(M) * col_size, 0);
    for (int i = 0; i < M; ++i)
      for (int k = 0; k < col_size; ++k) {
        size_t l = size_t(i) + size_t(k) * M;
        if (k < int(rows[i].size())) { val[l] = rows[i][k].second; colIdx[l] = rows[i][k].first; }
        else { val[l] = 0.0; colIdx[l] = i < N ? i : 0; }
      }
  }
}
double SparseMat::get(int i, int j) const {
  if (!ell) {
    for (int k = rowPtr[i]; k < rowPtr[i + 1]; ++k) if (colIdx[k] == j) return val[k];
    return 0.0;
  }
  for (int k = 0; k < col_size; ++k) {
    size_t l = size_t(i) + size_t(k) * M;
    if (colIdx[l] == j && val[l] != 0.0) return val[l];
  }
  return 0.0;
}
void SparseMat::matmul(const double* b, double* c) const {
  if (!ell) {
    for (int i = 0; i < M; ++i) {
      double s = 0.0;
      for (int k = rowPtr[i]; k < rowPtr[i + 1]; ++k) s += val[k] * b[colIdx[k]];
      c[i] = s;
    }
  } else {
    for (int i = 0; i < M; ++i) c[i] = 0.0;
    for (int k = 0; k < col_size; ++k)
      for (int i = 0; i < M; ++i) { size_t l = size_t(i) + size_t(k) * M; c[i] += val[l] * b[colIdx[l]]; }
  }
}
// sparsemat_matmul_CSR_1_2 (:476-512), sparsemat_matmul_ELL_1_2 (:587-634): c = A (b1 .* b2), evaluated as (A(j) * b1) * b2
void SparseMat::matmul_1_2(const double* b1, const double* b2, double* c) const {
  if (!ell) {
    for (int i = 0; i < M; ++i) {
      double s = 0.0;
      for (int k = rowPtr[i]; k < rowPtr[i + 1]; ++k) s = s + val[k] * b1[colIdx[k]] * b2[colIdx[k]];
      c[i] = s;
    }
  } else {
    for (int i = 0; i < M; ++i) c[i] = 0.0;
    for (int k = 0; k < col_size; ++k)
      for (int i = 0; i < M; ++i) { size_t l = size_t(i) + size_t(k) * M; c[i] = c[i] + val[l] * b1[colIdx[l]] * b2[colIdx[l]]; }
  }
}
// sparsemat_matmul_CSR_2 (:514-552), sparsemat_matmul_ELL_2 (:636-664): b(NQ,N), c(NQ,M) in Fortran order (q fastest)
void SparseMat::matmul2(const double* b, double* c, int NQ) const {
  for (size_t x = 0; x < size_t(NQ) * M; ++x) c[x] = 0.0;
  if (!ell) {
    for (int i = 0; i < M; ++i)
      for (int k = rowPtr[i]; k < rowPtr[i + 1]; ++k)
        for (int q = 0; q < NQ; ++q) c[q + size_t(NQ) * i] = c[q + size_t(NQ) * i] + val[k] * b[q + size_t(NQ) * colIdx[k]];
  } else {
    for (int k = 0; k < col_size; ++k)
      for (int i = 0; i < M; ++i) {
        size_t l = size_t(i) + size_t(k) * M;
        for (int q = 0; q < NQ; ++q) c[q + size_t(NQ) * i] = c[q + size_t(NQ) * i] + val[l] * b[q + size_t(NQ) * colIdx[l]];
      }
  }
}

}  // namespace feo
