// amg_refresh.cuh -- numeric re-setup of the multigrid hierarchy on the device (included by amg.cu).
//
// The symbolic part of a setup (aggregates, strength flags, the patterns of P, R = P^T, A P and R A P, the sliced-ELL
// layouts) is kept on the device after the host setup; when the coefficients of the level-0 matrix change on the same
// pattern (variable-density pressure equation of FractionalStepMultiphase, US/FractionalStepMultiphase.cpp:129-148; the
// momentum equation of every solver), the VALUES of every level are recomputed here by kernels, fp64, with the formulas
// of make_prolongator / build_hierarchy:
//   d, df, rho        diagonal, filtered diagonal, Gershgorin bounds            k_rf_diag
//   w = omegaS/rho/d  smoother weights                                          k_rf_weights
//   P = (I - omegaP/rhoP Df^-1 Af) T                                            k_rf_prolong
//   R = P^T           permutation of the entries of P                           k_rf_permute
//   A P, R (A P)      row-wise products on fixed patterns                       k_rf_product
//   A_c^-1            blocked Gauss-Jordan, 16 pivots per launch                k_rf_dense_fill/scatter, k_rf_gj_step
// Every output entry is owned by one lane, which walks the contributing products in a fixed order: no atomics, results
// do not depend on the launch geometry.
#pragma once

struct RfCsr {   // device CSR view
  const int *rp, *ci;
  const double *v;
};

__device__ __forceinline__ void rf_atomic_max(double *addr, double val) {   // val >= 0: bit patterns order like the values
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(val));
}

// level-0 CSR values out of the sliced-ELL slots of the Krylov matrix
__global__ void __launch_bounds__(256)
k_rf_gather(long long nnz, const int *__restrict__ src, const double *__restrict__ slotVals, double *__restrict__ v) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x)
    v[k] = slotVals[src[k]];
}

// scal[0] = max_i sum_k |a_ik| / |d_i|, scal[1] = max_i (|df_i| + sum_strong |a_ik|) / |df_i|, scal[3] = 1 on a zero diagonal
__global__ void __launch_bounds__(256)
k_rf_diag(int n, RfCsr A, const unsigned char *__restrict__ strong, double *__restrict__ diag, double *__restrict__ df,
          double *scal) {
  double m0 = 0., m1 = 0.;
  bool bad = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double d = 0., s = 0., f = 0., sp = 0.;
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k) {
      const double a = A.v[k];
      s += fabs(a);
      if (A.ci[k] == i) d += a;
      else if (strong) { if (strong[k]) sp += fabs(a); else f += a; }
    }
    diag[i] = d;
    if (d == 0.) { bad = true; continue; }
    m0 = fmax(m0, s / fabs(d));
    if (strong) {
      f += d;
      if (f == 0. || (f > 0.) != (d > 0.)) f = d;
      df[i] = f;
      m1 = fmax(m1, (fabs(f) + sp) / fabs(f));
    }
  }
  for (int o = 16; o; o >>= 1) {
    m0 = fmax(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (m0 > 0.) rf_atomic_max(scal + 0, m0);
    if (m1 > 0.) rf_atomic_max(scal + 1, m1);
  }
  if (bad) scal[3] = 1.;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_rf_weights(int n, const double *__restrict__ diag, const double *scal, double omegaS, T *__restrict__ w) {
  const double f = omegaS / scal[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) w[i] = (T)(f / diag[i]);
}

// sliced-ELL image (cycle precision) of a CSR matrix: slot -> CSR entry, -1 = padding
template <typename T>
__global__ void __launch_bounds__(256)
k_rf_fill(long long nSlots, const int *__restrict__ src, const double *__restrict__ v, T *__restrict__ out) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nSlots; k += (long long)gridDim.x * blockDim.x) {
    const int e = src[k];
    out[k] = e >= 0 ? (T)v[e] : (T)0;
  }
}

// P(i, J) = [J == agg(i)] (1 - w) - w / df_i  sum over the strong k of row i with agg(col_k) == J of a_ik,  w = omegaP / rhoP
template <int LANES>
__global__ void __launch_bounds__(256)
k_rf_prolong(int n, RfCsr A, const unsigned char *__restrict__ strong, const int *__restrict__ agg,
             const double *__restrict__ df, const double *scal, double omegaP, const int *__restrict__ pRp,
             const int *__restrict__ pCi, double *__restrict__ pV) {
  const double w = omegaP / scal[1];
  const int lane = threadIdx.x % LANES;
  for (long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / LANES; g < n;
       g += (long long)gridDim.x * blockDim.x / LANES) {
    const int i = (int)g, a0 = A.rp[i], a1 = A.rp[i + 1], own = agg[i];
    const double f = -w / df[i];
    for (int e = pRp[i] + lane; e < pRp[i + 1]; e += LANES) {
      const int J = pCi[e];
      double acc = 0.;
      for (int k = a0; k < a1; ++k)
        if (strong[k] && agg[A.ci[k]] == J) acc += f * A.v[k];
      pV[e] = (J == own ? 1. - w : 0.) + acc;
    }
  }
}

__global__ void __launch_bounds__(256)
k_rf_permute(long long nnz, const int *__restrict__ src, const double *__restrict__ in, double *__restrict__ out) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x)
    out[k] = in[src[k]];
}

// C = A B on the fixed pattern of C: LANES lanes per row, lane e owns entry e of the row and adds up the products
// a_ik b_kJ whose column is its own, k in row order, the entries of row k of B in their order
template <int LANES>
__global__ void __launch_bounds__(256)
k_rf_product(int n, RfCsr A, RfCsr B, const int *__restrict__ cRp, const int *__restrict__ cCi, double *__restrict__ cV) {
  const int lane = threadIdx.x % LANES;
  for (long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / LANES; g < n;
       g += (long long)gridDim.x * blockDim.x / LANES) {
    const int i = (int)g, a0 = A.rp[i], a1 = A.rp[i + 1];
    for (int e = cRp[i] + lane; e < cRp[i + 1]; e += LANES) {
      const int J = cCi[e];
      double acc = 0.;
      for (int k = a0; k < a1; ++k) {
        const int r = A.ci[k];
        const double a = A.v[k];
        for (int q = B.rp[r]; q < B.rp[r + 1]; ++q)
          if (B.ci[q] == J) acc += a * B.v[q];
      }
      cV[e] = acc;
    }
  }
}

// ---- dense inverse of the coarsest operator
__global__ void __launch_bounds__(1024)
k_rf_diag_mean(int n, const double *__restrict__ diag, double *scal) {   // one block: scal[2] = sum(diag) / n, fixed order
  __shared__ double sh[1024];
  double s = 0.;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += diag[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) scal[2] = sh[0] / n;
}

__global__ void __launch_bounds__(256)
k_rf_dense_fill(int n, const double *scal, int singular, double *__restrict__ M) {
  const double c = singular ? scal[2] / n : 0.;   // + (mean diag / n) 1 1^T: the constant leaves the null space
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < (long long)n * n; k += (long long)gridDim.x * blockDim.x)
    M[k] = c;
}

__global__ void __launch_bounds__(256)
k_rf_dense_scatter(int n, RfCsr A, double *__restrict__ M) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k) M[(size_t)i * n + A.ci[k]] += A.v[k];
}

// One step of the blocked in-place Gauss-Jordan inversion (no pivoting: the coarsest operators are definite or
// diagonally dominant; a vanishing pivot raises flag[0] and the caller falls back to the host setup).  K = the pivot
// rows/columns k0 .. k0 + b - 1, B = src(K, K)^-1:
//   dst(K, K) = B,  dst(K, c) = B src(K, c),  dst(r, K) = -src(r, K) B,  dst(r, c) = src(r, c) - src(r, K) B src(K, c)
// Every CTA inverts the small pivot block itself (16 x 16 in shared memory) and updates one 64 x 64 tile.
constexpr int kGjBlock = 16;
constexpr int kGjTile = 64;
__global__ void __launch_bounds__(256)
k_rf_gj_step(int n, int k0, const double *__restrict__ src, double *__restrict__ dst, int *flag) {
  __shared__ double B[kGjBlock][kGjBlock + 1], L[kGjTile][kGjBlock + 1], U[kGjBlock][kGjTile + 1], LB[kGjTile][kGjBlock + 1];
  const int b = min(kGjBlock, n - k0), t = threadIdx.x;
  const int r0 = blockIdx.y * kGjTile, c0 = blockIdx.x * kGjTile;
  {  // pivot block and its inverse (unblocked Gauss-Jordan, thread (i, j) owns entry (i, j))
    const int i = t / kGjBlock, j = t % kGjBlock;
    const bool in = i < b && j < b;
    if (in) B[i][j] = src[(size_t)(k0 + i) * n + k0 + j];
    __syncthreads();
    for (int k = 0; k < b; ++k) {
      const double p = B[k][k];
      if (t == 0 && !(fabs(p) > 0.) ) atomicExch(flag, 1);
      double val = 0.;
      if (in) {
        const double ip = 1. / p;
        if (i == k && j == k) val = ip;
        else if (i == k) val = B[k][j] * ip;
        else if (j == k) val = -B[i][k] * ip;
        else val = B[i][j] - B[i][k] * B[k][j] * ip;
      }
      __syncthreads();
      if (in) B[i][j] = val;
      __syncthreads();
    }
  }
  // panels of this tile: L = src(rows, K), U = src(K, cols)
  for (int e = t; e < kGjTile * kGjBlock; e += 256) {
    const int r = e / kGjBlock, j = e % kGjBlock;
    L[r][j] = (r0 + r < n && j < b) ? src[(size_t)(r0 + r) * n + k0 + j] : 0.;
    const int jj = e / kGjTile, c = e % kGjTile;
    U[jj][c] = (c0 + c < n && jj < b) ? src[(size_t)(k0 + jj) * n + c0 + c] : 0.;
  }
  __syncthreads();
  for (int e = t; e < kGjTile * kGjBlock; e += 256) {
    const int r = e / kGjBlock, j = e % kGjBlock;
    double acc = 0.;
    for (int i = 0; i < b; ++i) acc += L[r][i] * B[i][j];
    LB[r][j] = acc;
  }
  __syncthreads();
  for (int e = t; e < kGjTile * kGjTile; e += 256) {
    const int r = e / kGjTile, c = e % kGjTile;
    const int gr = r0 + r, gc = c0 + c;
    if (gr >= n || gc >= n) continue;
    const bool rk = gr >= k0 && gr < k0 + b, ck = gc >= k0 && gc < k0 + b;
    double val;
    if (rk && ck) val = B[gr - k0][gc - k0];
    else if (ck) val = -LB[r][gc - k0];
    else if (rk) {
      val = 0.;
      for (int j = 0; j < b; ++j) val += B[gr - k0][j] * U[j][c];
    } else {
      val = src[(size_t)gr * n + gc];
      for (int j = 0; j < b; ++j) val -= LB[r][j] * U[j][c];
    }
    dst[(size_t)gr * n + gc] = val;
  }
}
