// Marching fast path of the pressure (SIPG Helmholtz) operator for fully periodic 3-D boxes
// (apply_add with NS_stage == 2, NS.cc:1407-1414; cell / face forms NS.cc:1230-1292).  DESIGN.md
// "marching kernel" explains the derivation; in short
//
//   A u = (Mx (x) My (x) Mz) [ kappa u + sum_d ( Kt'_d u  +  sum_s cG_d,s G'_s + cV_d,s V'_s ) ]
//
// where Kt'_d is the 1-D operator M^-1 K / h_d^2 with the cell's OWN share of both SIPG face fluxes
// folded in, and (V', G') are the value and outward normal derivative of the face neighbour on the
// shared face (only 2 numbers per face point cross a face).  One CTA owns a tile of 8 x 4 cells and
// marches along z.  Warps are specialised and software-pipelined over the planes of the march:
//
//   Z warps   column threads: stream the dofs of plane p+1 from HBM (coalesced), do all z-direction work
//             in registers (Kt'_z, z-face terms: the traces of the plane below/above are carried in
//             registers, they never touch shared memory), hand u and the partial sum to the XY warps,
//             and later apply Mz to the result and stream it to HBM (coalesced).
//   XY warps  one thread per (cell, z-layer): holds an n x n layer in registers, does the x and y
//             sweeps (Kt'_x, Kt'_y, Mx, My) without shared-memory round trips; only face traces are
//             exchanged between threads.
//   H warps   read the x / y halo cells of the tile and reduce them to (V', G') traces.
//
// Hand-offs use named barriers (bar.arrive / bar.sync), stages are double buffered.
#include "pressure_op.cuh"

namespace dgx {
namespace {

constexpr int TXC = 8, TYC = 4, TC = TXC * TYC;   // cells per plane of the tile = lanes of a warp
static_assert(TC == 32, "one lane per cell of the plane");

template <typename T, int n>
struct FastParams {
  T Kt[3][n * n];   // folded 1-D operator, [d][i*n+j]
  T Mh[3][n * n];   // h_d M
  T g[2][n];        // own outward normal derivative weights n_s l_j'(s) (reference scaling)
  T cG[3][2][n];    // +1/2 a_d,s
  T cV[3][2][n];    // -C a_d,s + 1/2 b_d,s
  T kappa;
};

struct FastGeom {
  int L;                 // refinement level of the cells
  int c0x, c0y;          // coarse cells per direction (x, y)
  int lox, loy, loz;     // owned box origin (cell coordinates)
  int ntx, nty, nsz;     // tiles in x, y; z segments
  int NZ;                // planes per segment
  long long first_cell;  // first owned cell (hierarchical index)
  int n_owned;
};

template <int n>
struct Roles {
  static constexpr int NXY = n;                       // XY warps: one per z-layer
  static constexpr int NZW = (n == 5) ? 5 : 4;        // column warps
  static constexpr int NH = 2;                        // halo warps
  static constexpr int NW = NXY + NZW + NH;
  static constexpr int NCOL = (n * n + NZW - 1) / NZW;  // column units (32 columns each) per Z warp
};

// barrier ids (0 is __syncthreads)
enum { B_FULL = 1, B_UFREE = 3, B_QREADY = 5, B_XY = 7 };

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
template <typename T> __device__ __forceinline__ void keep(T &v);
template <> __device__ __forceinline__ void keep<double>(double &v) { asm volatile("" : "+d"(v)::"memory"); }
template <> __device__ __forceinline__ void keep<float>(float &v) { asm volatile("" : "+f"(v)::"memory"); }

__device__ __forceinline__ unsigned part1by2(unsigned x) {
  x &= 0x3ffu;
  x = (x | (x << 16)) & 0x030000ffu;
  x = (x | (x << 8)) & 0x0300f00fu;
  x = (x | (x << 4)) & 0x030c30c3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}

// hierarchical (coarse lexicographic, Morton inside) index of cell (X, Y, Z), local to the owned range
__device__ __forceinline__ int cell_index(const FastGeom &G, int X, int Y, int Z) {
  const int mask = (1 << G.L) - 1;
  const long long coarse = (X >> G.L) + (long long)G.c0x * ((Y >> G.L) + (long long)G.c0y * (Z >> G.L));
  const unsigned loc = part1by2(X & mask) | (part1by2(Y & mask) << 1) | (part1by2(Z & mask) << 2);
  return (int)((coarse << (3 * G.L)) + loc - G.first_cell);
}

template <int n, typename T>
__global__ void __launch_bounds__(Roles<n>::NW * 32, 1)
pressure_march_kernel(const __grid_constant__ FastParams<T, n> P, const __grid_constant__ FastGeom G,
                      const int *__restrict__ nbr, const T *__restrict__ src, T *__restrict__ dst, int add) {
  using R = Roles<n>;
  constexpr int NL = n * n, ND = NL * n;
  constexpr int CNT_FULL = (R::NXY + R::NZW + R::NH) * 32, CNT_UFREE = (R::NXY + R::NZW) * 32,
                CNT_QREADY = (R::NXY + R::NZW + R::NH) * 32, CNT_XY = R::NXY * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *U = reinterpret_cast<T *>(smem_raw);          // [2][TC][ND]
  T *W = U + 2 * TC * ND;                          // [2][TC][ND]
  T *TX = W + 2 * TC * ND;                         // [2 side][NL][TC]   own outward-normal-derivative traces, x faces
  T *TY = TX + 2 * NL * TC;                        // [2 side][NL][TC]
  T *HX = TY + 2 * NL * TC;                        // [2 buf][2 side][TYC][NL][2]  (V', G') of the x halo
  T *HY = HX + 2 * 2 * TYC * NL * 2;               // [2 buf][2 side][TXC][NL][2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // role assignment balanced over the 4 sub-partitions (warp % 4): XY warps carry twice the work
  int role, ridx;   // 0 = XY, 1 = Z, 2 = H
  if (n == 5) {
    const int role_tab[12] = {0, 0, 0, 1, 0, 0, 1, 1, 2, 2, 1, 1};
    const int idx_tab[12] = {0, 2, 4, 2, 1, 3, 0, 3, 0, 1, 1, 4};
    role = role_tab[warp]; ridx = idx_tab[warp];
  } else {
    role = warp < R::NXY ? 0 : (warp < R::NXY + R::NZW ? 1 : 2);
    ridx = warp < R::NXY ? warp : (warp < R::NXY + R::NZW ? warp - R::NXY : warp - R::NXY - R::NZW);
  }

  const int job = blockIdx.x;
  const int tx = job % G.ntx, ty = (job / G.ntx) % G.nty, sz = job / (G.ntx * G.nty);
  const int x0 = G.lox + tx * TXC, y0 = G.loy + ty * TYC, z0 = G.loz + sz * G.NZ;
  const int NZ = G.NZ;
  const size_t no = (size_t)G.n_owned;

  if (role == 1) {
    // ------------------------------------------------------------------ Z warps (column threads)
    const int zw = ridx;
    int ccell[R::NCOL], cij[R::NCOL], cxy[R::NCOL];
    bool cval[R::NCOL];
#pragma unroll
    for (int m = 0; m < R::NCOL; ++m) {
      const int unit = zw + R::NZW * m;
      cval[m] = unit < NL;
      const int id = (cval[m] ? unit : 0) * 32 + lane;
      ccell[m] = id / NL;
      cij[m] = id - ccell[m] * NL;
      cxy[m] = cell_index(G, x0 + (ccell[m] & (TXC - 1)), y0 + ccell[m] / TXC, z0);   // index at plane z0
    }
    T r[R::NCOL][n], wz[R::NCOL][n], Gt[R::NCOL], Vt[R::NCOL];
    // z-dependent part of the index relative to plane z0 (tile lies inside one coarse cell in x, y)
    auto plane_shift = [&](int z) -> long long {
      const int mask = (1 << G.L) - 1;
      const long long c1 = ((long long)G.c0x * G.c0y * (z >> G.L) << (3 * G.L)) + ((long long)part1by2(z & mask) << 2);
      const long long c0 = ((long long)G.c0x * G.c0y * (z0 >> G.L) << (3 * G.L)) + ((long long)part1by2(z0 & mask) << 2);
      return c1 - c0;
    };
    auto load_plane = [&](int q) {   // q in -1 .. NZ : plane z0 + q (ghost planes through the neighbour table)
#pragma unroll
      for (int m = 0; m < R::NCOL; ++m) {
        if (!cval[m]) continue;
        long long ci;
        if (q < 0) ci = nbr[4 * no + cxy[m]];
        else if (q >= NZ) ci = nbr[5 * no + (cxy[m] + plane_shift(z0 + NZ - 1))];
        else ci = cxy[m] + plane_shift(z0 + q);
        const T *p = src + (size_t)ci * ND + cij[m];
#pragma unroll
        for (int k = 0; k < n; ++k) r[m][k] = __ldg(p + k * NL);
      }
    };
    auto top_traces = [&]() {
#pragma unroll
      for (int m = 0; m < R::NCOL; ++m) {
        T gsum = 0;
#pragma unroll
        for (int k = 0; k < n; ++k) gsum += P.g[1][k] * r[m][k];
        Gt[m] = gsum; Vt[m] = r[m][n - 1];
      }
    };
    auto apply_mz_store = [&](int p) {   // Z2: dst(plane p) = Mz q
      const T *Ws = W + (size_t)(p & 1) * TC * ND;
      const long long sh = plane_shift(z0 + p);
#pragma unroll
      for (int m = 0; m < R::NCOL; ++m) {
        if (!cval[m]) continue;
        T qv[n];
#pragma unroll
        for (int k = 0; k < n; ++k) qv[k] = Ws[ccell[m] * ND + cij[m] + k * NL];
        T *o = dst + (size_t)(cxy[m] + sh) * ND + cij[m];
#pragma unroll
        for (int k = 0; k < n; ++k) {
          T acc = 0;
#pragma unroll
          for (int kk = 0; kk < n; ++kk) acc += P.Mh[2][k * n + kk] * qv[kk];
          o[k * NL] = add ? o[k * NL] + acc : acc;
        }
      }
    };

    load_plane(-1);
    top_traces();
    load_plane(0);
    for (int p = -1; p < NZ; ++p) {
      const int q = p + 1;   // plane currently held in r
      if (p >= 0) {
        // finalise plane p with the bottom traces of plane q, publish the z-partial sum
        T *Ws = W + (size_t)(p & 1) * TC * ND;
#pragma unroll
        for (int m = 0; m < R::NCOL; ++m) {
          if (!cval[m]) continue;
          T gb = 0;
#pragma unroll
          for (int k = 0; k < n; ++k) gb += P.g[0][k] * r[m][k];
          const T vb = r[m][0];
#pragma unroll
          for (int k = 0; k < n; ++k) Ws[ccell[m] * ND + cij[m] + k * NL] = wz[m][k] + P.cG[2][1][k] * gb + P.cV[2][1][k] * vb;
        }
        bar_arrive(B_FULL + (p & 1), CNT_FULL);
      }
      if (q < NZ) {
#pragma unroll
        for (int m = 0; m < R::NCOL; ++m) {
#pragma unroll
          for (int k = 0; k < n; ++k) {
            T acc = P.kappa * r[m][k] + P.cG[2][0][k] * Gt[m] + P.cV[2][0][k] * Vt[m];
#pragma unroll
            for (int kk = 0; kk < n; ++kk) acc += P.Kt[2][k * n + kk] * r[m][kk];
            wz[m][k] = acc;
          }
        }
        top_traces();
        if (q >= 2) bar_sync(B_UFREE + (q & 1), CNT_UFREE);
        T *Us = U + (size_t)(q & 1) * TC * ND;
#pragma unroll
        for (int m = 0; m < R::NCOL; ++m) {
          if (!cval[m]) continue;
#pragma unroll
          for (int k = 0; k < n; ++k) Us[ccell[m] * ND + cij[m] + k * NL] = r[m][k];
        }
        load_plane(q + 1);
      }
      if (p >= 1) {
        bar_sync(B_QREADY + ((p - 1) & 1), CNT_QREADY);
        apply_mz_store(p - 1);
      }
    }
    // drain: the UFREE arrivals of the last two planes, then the last result plane
    if (NZ >= 2) bar_sync(B_UFREE + ((NZ - 2) & 1), CNT_UFREE);
    bar_sync(B_UFREE + ((NZ - 1) & 1), CNT_UFREE);
    bar_sync(B_QREADY + ((NZ - 1) & 1), CNT_QREADY);
    apply_mz_store(NZ - 1);
  } else if (role == 0) {
    // ------------------------------------------------------------------ XY warps (layer threads)
    const int k = ridx, c = lane, cx = c & (TXC - 1), cy = c / TXC;
    for (int p = 0; p < NZ; ++p) {
      const int st = p & 1;
      bar_sync(B_FULL + st, CNT_FULL);
      const T *Uc = U + (size_t)st * TC * ND;
      T *Wp = W + (size_t)st * TC * ND + c * ND + k * NL;
      const T *Up = Uc + c * ND + k * NL;
      T u[n][n], acc[n][n];
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int i = 0; i < n; ++i) { u[j][i] = Up[j * n + i]; acc[j][i] = Wp[j * n + i]; }
      // own traces (outward normal derivative, reference scaling) on the 4 lateral faces of the layer
#pragma unroll
      for (int j = 0; j < n; ++j) {
        T g0 = 0, g1 = 0;
#pragma unroll
        for (int i = 0; i < n; ++i) { g0 += P.g[0][i] * u[j][i]; g1 += P.g[1][i] * u[j][i]; }
        TX[(0 * NL + j + n * k) * TC + c] = g0;
        TX[(1 * NL + j + n * k) * TC + c] = g1;
      }
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T g0 = 0, g1 = 0;
#pragma unroll
        for (int j = 0; j < n; ++j) { g0 += P.g[0][j] * u[j][i]; g1 += P.g[1][j] * u[j][i]; }
        TY[(0 * NL + i + n * k) * TC + c] = g0;
        TY[(1 * NL + i + n * k) * TC + c] = g1;
      }
      bar_sync(B_XY, CNT_XY);
      // neighbour traces: x faces
      const T *hx = HX + (size_t)st * (2 * TYC * NL * 2);
      const T *hy = HY + (size_t)st * (2 * TXC * NL * 2);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const bool inside = s ? (cx < TXC - 1) : (cx > 0);
        const int cn = s ? c + 1 : c - 1;
#pragma unroll
        for (int j = 0; j < n; ++j) {
          const T *h = hx + ((s * TYC + cy) * NL + j + n * k) * 2;
          const T gn = *(inside ? &TX[((1 - s) * NL + j + n * k) * TC + cn] : h + 1);
          const T vn = *(inside ? &Uc[cn * ND + k * NL + j * n + (1 - s) * (n - 1)] : h);
#pragma unroll
          for (int i = 0; i < n; ++i) acc[j][i] += P.cG[0][s][i] * gn + P.cV[0][s][i] * vn;
        }
      }
      // y faces
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const bool inside = s ? (cy < TYC - 1) : (cy > 0);
        const int cn = s ? c + TXC : c - TXC;
#pragma unroll
        for (int i = 0; i < n; ++i) {
          const T *h = hy + ((s * TXC + cx) * NL + i + n * k) * 2;
          const T gn = *(inside ? &TY[((1 - s) * NL + i + n * k) * TC + cn] : h + 1);
          const T vn = *(inside ? &Uc[cn * ND + k * NL + (1 - s) * (n - 1) * n + i] : h);
#pragma unroll
          for (int j = 0; j < n; ++j) acc[j][i] += P.cG[1][s][j] * gn + P.cV[1][s][j] * vn;
        }
      }
      // all shared-memory reads of this plane's u and traces are done: make sure they have landed
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int i = 0; i < n; ++i) { keep(u[j][i]); keep(acc[j][i]); }
      bar_arrive(B_UFREE + st, CNT_UFREE);
      bar_sync(B_XY, CNT_XY);   // nobody overwrites TX / TY before every layer thread has read them
      // folded 1-D operators in x and y
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int i = 0; i < n; ++i) {
          T a = acc[j][i];
#pragma unroll
          for (int l = 0; l < n; ++l) a += P.Kt[0][i * n + l] * u[j][l];
#pragma unroll
          for (int l = 0; l < n; ++l) a += P.Kt[1][j * n + l] * u[l][i];
          acc[j][i] = a;
        }
      // mass in x, then y (u is dead: reuse it as the temporary)
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int i = 0; i < n; ++i) {
          T a = 0;
#pragma unroll
          for (int l = 0; l < n; ++l) a += P.Mh[0][i * n + l] * acc[j][l];
          u[j][i] = a;
        }
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int i = 0; i < n; ++i) {
          T a = 0;
#pragma unroll
          for (int l = 0; l < n; ++l) a += P.Mh[1][j * n + l] * u[l][i];
          Wp[j * n + i] = a;
        }
      bar_arrive(B_QREADY + st, CNT_QREADY);
    }
  } else {
    // ------------------------------------------------------------------ H warps (halo traces)
    constexpr int NI = (2 * TYC + 2 * TXC) * NL, NT = R::NH * 32, NIT = (NI + NT - 1) / NT;
    const int t = ridx * 32 + lane;
    // static description of this thread's items
    int hcell[NIT], hoff[NIT], hstr[NIT], hface[NIT], hslot[NIT];
    bool hval[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int item = t + it * NT;
      hval[it] = item < NI;
      const int e = hval[it] ? item : 0;
      if (e < 2 * TYC * NL) {            // x halo: side s, row cy, face point fp = j + n k  -> line along i
        const int s = e / (TYC * NL), rem = e - s * TYC * NL, cy = rem / NL, fp = rem - cy * NL;
        hcell[it] = (s ? TXC - 1 : 0) + TXC * cy;
        hface[it] = s;
        hoff[it] = fp * n; hstr[it] = 1;
        hslot[it] = ((s * TYC + cy) * NL + fp) * 2;                      // into HX
      } else {                            // y halo: side s, column cx, face point fp = i + n k -> line along j
        const int e2 = e - 2 * TYC * NL;
        const int s = e2 / (TXC * NL), rem = e2 - s * TXC * NL, cx = rem / NL, fp = rem - cx * NL;
        hcell[it] = cx + TXC * (s ? TYC - 1 : 0);
        hface[it] = 2 + s;
        hoff[it] = (fp % n) + (fp / n) * NL; hstr[it] = n;
        hslot[it] = 2 * 2 * TYC * NL * 2 + ((s * TXC + cx) * NL + fp) * 2;   // into HY (relative to HX base, buffer 0)
      }
      hcell[it] = cell_index(G, x0 + (hcell[it] & (TXC - 1)), y0 + hcell[it] / TXC, z0);
    }
    const int mask = (1 << G.L) - 1;
    const long long sh0 = ((long long)G.c0x * G.c0y * (z0 >> G.L) << (3 * G.L)) + ((long long)part1by2(z0 & mask) << 2);
    for (int p = 0; p < NZ + 2; ++p) {
      T hv[NIT], hg[NIT];
      if (p < NZ) {
        const int z = z0 + p;
        const long long sh = ((long long)G.c0x * G.c0y * (z >> G.L) << (3 * G.L)) + ((long long)part1by2(z & mask) << 2) - sh0;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          hv[it] = 0; hg[it] = 0;
          if (!hval[it]) continue;
          const int s = hface[it] & 1;
          const long long hc = nbr[(size_t)hface[it] * no + (hcell[it] + sh)];
          const T *ptr = src + (size_t)hc * ND + hoff[it];
          T gsum = 0, last = 0;
#pragma unroll
          for (int l = 0; l < n; ++l) {
            const T x = __ldg(ptr + l * hstr[it]);
            gsum += (s ? P.g[0][l] : P.g[1][l]) * x;    // neighbour's facing side is 1 - s
            if (l == (s ? 0 : n - 1)) last = x;
          }
          hv[it] = last; hg[it] = gsum;
        }
      }
      if (p >= 2) bar_sync(B_QREADY + (p & 1), CNT_QREADY);
      if (p < NZ) {
        const int st = p & 1;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          if (!hval[it]) continue;
          // HX buffers are [2][...], HY buffers follow; hslot for HY was computed relative to the HX base
          T *base = (hface[it] < 2) ? (HX + (size_t)st * (2 * TYC * NL * 2) + hslot[it])
                                    : (HY + (size_t)st * (2 * TXC * NL * 2) + (hslot[it] - 2 * 2 * TYC * NL * 2));
          base[0] = hv[it]; base[1] = hg[it];
        }
        bar_arrive(B_FULL + st, CNT_FULL);
      }
    }
  }
}

template <typename T, int n>
void fill_fast_params(FastParams<T, n> &F, const dgx_op *op) {
  const Shape1D &s = get_shape(n - 1, n);
  const double C = (double)n * n;
  for (int d = 0; d < 3; ++d) {
    const double h = op->mf->h[d], ih2 = 1.0 / (h * h);
    double Kt[n * n];
    for (int i = 0; i < n * n; ++i) Kt[i] = ih2 * s.Kt[i];
    for (int sd = 0; sd < 2; ++sd) {
      const double ns = sd ? 1.0 : -1.0;
      const int e = sd * (n - 1);
      for (int i = 0; i < n; ++i) {
        const double a = ih2 * s.a[sd][i], b = ih2 * ns * s.b[sd][i];
        for (int j = 0; j < n; ++j) Kt[i * n + j] += a * (-0.5 * ns * s.g[sd][j] + (j == e ? C : 0.0)) - (j == e ? 0.5 * b : 0.0);
        F.cG[d][sd][i] = (T)(0.5 * a);
        F.cV[d][sd][i] = (T)(-C * a + 0.5 * b);
      }
    }
    for (int i = 0; i < n * n; ++i) { F.Kt[d][i] = (T)Kt[i]; F.Mh[d][i] = (T)(h * s.M[i]); }
  }
  for (int sd = 0; sd < 2; ++sd)
    for (int j = 0; j < n; ++j) F.g[sd][j] = (T)((sd ? 1.0 : -1.0) * s.g[sd][j]);
  F.kappa = (T)pressure_kappa(op);
}

template <int n, typename T>
int launch_march(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, const FastGeom &G) {
  using R = Roles<n>;
  constexpr int NL = n * n, ND = NL * n;
  FastParams<T, n> F;
  fill_fast_params<T, n>(F, op);
  const size_t smem = sizeof(T) * (size_t)(4 * TC * ND + 4 * NL * TC + 2 * 2 * TYC * NL * 2 + 2 * 2 * TXC * NL * 2);
  static bool configured = false;
  if (!configured) {
    DGX_CUDA_CHECK(cudaFuncSetAttribute(pressure_march_kernel<n, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int grid = G.ntx * G.nty * G.nsz;
  pressure_march_kernel<n, T><<<grid, R::NW * 32, smem, op->mf->ctx->stream>>>(F, G, op->mf->d_nbr, (const T *)src->d, (T *)dst->d, add ? 1 : 0);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

}  // namespace

// The marching kernel applies when every face is interior (fully periodic mesh) and the owned cells form a
// box of the hierarchical order whose extents are multiples of the tile.
int pressure_apply_fast(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add) {
  dgx_mf *mf = op->mf;
  if (mf->dim != 3 || !mf->all_interior || !mf->box_ok) return DGX_ERR_UNSUPPORTED;
  if (op->degree != 4 && op->degree != 3) return DGX_ERR_UNSUPPORTED;
  const Mesh &m = mf->mesh->m;
  FastGeom G;
  G.L = mf->level;
  G.c0x = m.cells0[0]; G.c0y = m.cells0[1];
  G.lox = mf->box_lo[0]; G.loy = mf->box_lo[1]; G.loz = mf->box_lo[2];
  const int EX = mf->box_ext[0], EY = mf->box_ext[1], EZ = mf->box_ext[2];
  if (G.L < 3 || EX % TXC || EY % TYC || EZ < 2) return DGX_ERR_UNSUPPORTED;
  G.ntx = EX / TXC; G.nty = EY / TYC;
  int NZ = EZ;
  while (NZ > 64 && NZ % 2 == 0) NZ /= 2;
  G.NZ = NZ; G.nsz = EZ / NZ;
  G.first_cell = mf->first_cell;
  G.n_owned = (int)mf->n_owned;
  if (mf->dtype == DGX_F64) return op->degree == 4 ? launch_march<5, double>(op, dst, src, add, G) : launch_march<4, double>(op, dst, src, add, G);
  return op->degree == 4 ? launch_march<5, float>(op, dst, src, add, G) : launch_march<4, float>(op, dst, src, add, G);
}

}  // namespace dgx
