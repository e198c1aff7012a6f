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
#ifndef DGX_MB4F
#define DGX_MB4F 2
#endif
#ifndef DGX_NZW4
#define DGX_NZW4 8
#endif
#ifndef DGX_XY_UNROLLED
#define DGX_XY_UNROLLED 1
#endif
#include <type_traits>
#include <utility>
#include <vector>

#include <cmath>
#include <cstdlib>
#include <cstring>

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
  int wait_mode;         // v2 kernel: bit 0 / 1 / 2 = XY / Z / H warps spin on their mbarriers instead of sleeping
};

template <int n>
struct Roles {
  static constexpr int NXY = n;                       // XY warps: one per z-layer
  static constexpr int NZW = (n == 5) ? 5 : DGX_NZW4; // column warps
  static constexpr int NH = 2;                        // halo warps
  static constexpr int NW = NXY + NZW + NH;
  static constexpr int NCOL = (n * n + NZW - 1) / NZW;  // column units (32 columns each) per Z warp
  static constexpr int NSTAGE = 3;                    // stages of the u planes
};
// resident CTAs per SM the register allocation is sized for (shared memory allows as many)
template <int n, typename T> struct MinBlocks { static constexpr int value = 1; };
template <> struct MinBlocks<5, float> { static constexpr int value = 2; };
// (4, double): shared memory allows one CTA per SM anyway, so take the registers
template <> struct MinBlocks<4, float> { static constexpr int value = DGX_MB4F; };
template <> struct MinBlocks<3, float> { static constexpr int value = 3; };
template <> struct MinBlocks<3, double> { static constexpr int value = 2; };

// barrier ids (0 is __syncthreads)
enum { B_FULL = 1, B_QREADY = 3, B_XY = 5 };

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
template <typename T> __device__ __forceinline__ void keep(T &v);
template <> __device__ __forceinline__ void keep<double>(double &v) { asm volatile("" : "+d"(v)::"memory"); }
template <> __device__ __forceinline__ void keep<float>(float &v) { asm volatile("" : "+f"(v)::"memory"); }

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  unsigned ok = 0;
  for (int spin = 0; !ok; ++spin) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (spin > (1 << 24)) __trap();   // never hang the device on a lost copy
  }
}
// 1-D bulk async copy global -> shared, completion counted on an mbarrier (TMA engine, SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

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
// z-dependent part of the index (the tile lies inside one coarse cell in x and y)
__device__ __forceinline__ long long plane_term(const FastGeom &G, int z) {
  const int mask = (1 << G.L) - 1;
  return ((long long)G.c0x * G.c0y * (z >> G.L) << (3 * G.L)) + ((long long)part1by2(z & mask) << 2);
}

// Cells of a tile plane are kept in shared memory in their memory (Morton) order, so that a plane is 8 bulk
// copies of 4 consecutive cells: slot bits (x0, y0, x1, y1, x2).  One lane of an XY warp per slot.
__device__ __forceinline__ int slot_cx(int s) { return (s & 1) | ((s >> 1) & 2) | ((s >> 2) & 4); }
__device__ __forceinline__ int slot_cy(int s) { return ((s >> 1) & 1) | ((s >> 2) & 2); }
__device__ __forceinline__ int slot_of(int cx, int cy) { return (cx & 1) | ((cy & 1) << 1) | ((cx & 2) << 1) | ((cy & 2) << 2) | ((cx & 4) << 2); }

template <int n, typename T, bool ADD>
__global__ void __launch_bounds__(Roles<n>::NW * 32, MinBlocks<n, T>::value)
pressure_march_kernel(const __grid_constant__ FastParams<T, n> P, const __grid_constant__ FastGeom G,
                      const int *__restrict__ nbr, const T *__restrict__ src, T *__restrict__ dst, const int *__restrict__ jobs,
                      const __grid_constant__ Epilogue<T> E) {
  using R = Roles<n>;
  constexpr int NL = n * n, ND = NL * n, NS = R::NSTAGE;
  constexpr int CNT_FULL = (R::NXY + R::NZW + R::NH) * 32,
                CNT_QREADY = (R::NXY + R::NZW + R::NH) * 32, CNT_XY = R::NXY * 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // For even n the cell stride n^3 maps every lane (= cell) of an XY warp to the same bank: those instantiations keep a
  // second, padded copy of u (written by the Z warps, which hold the columns in registers anyway) and a padded W.
  constexpr bool PAD = (n % 2 == 0);
  // straight-line XY sweeps where the register budget allows it (fp64: one CTA per SM, 168 registers); rolled streaming
  // loops for fp32, which runs 2-3 CTAs per SM at 64-96 registers
  constexpr bool XY_UNROLLED = DGX_XY_UNROLLED && sizeof(T) == 8;
  constexpr int NDP = PAD ? ND + 1 : ND;
  T *U = reinterpret_cast<T *>(smem_raw);          // [NS][TC][ND]   u planes (bulk-copied, memory order)
  T *W = U + NS * TC * ND;                         // [2][TC][NDP]   z-partial sums, then the xy result
  T *UP = W + 2 * TC * NDP;                        // [2][TC][NDP]   padded u planes (PAD only)
  T *TX = UP + (PAD ? 2 * TC * NDP : 0);           // [2 side][NL][TC]   own outward-normal-derivative traces, x faces
  T *TY = TX + 2 * NL * TC;                        // [2 side][NL][TC]
  T *HXY = TY + 2 * NL * TC;                       // [2 buf]{[2 side][TYC][NL][2], [2 side][TXC][NL][2]}  (V', G') of the x / y halo
  constexpr int HBUF = (2 * TYC + 2 * TXC) * NL * 2;
  // fused update, padded instantiations: the planes of rhs, d and xold that the Z warps need when they store a plane are
  // bulk-copied one plane ahead into AUX (the loads would otherwise sit on the critical path of the producer warps)
  const bool AUXON = PAD && E.on;
  T *AUX = HXY + 2 * HBUF;                         // [3][TC][ND]  rhs, d, xold of the plane being stored (fused + PAD only)
  unsigned long long *ubar = reinterpret_cast<unsigned long long *>(AUX + (AUXON ? 3 * TC * ND : 0));   // [NS]
  int *ucnt = reinterpret_cast<int *>(ubar + NS);   // [NS] warps that have finished reading the plane held by each stage
  unsigned long long *abar = ubar + NS + (NS * 4 + 7) / 8;   // behind ucnt, 8-byte aligned
  int *acnt = reinterpret_cast<int *>(abar + 1);

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
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&ubar[s], 1);
    for (int s = 0; s < NS; ++s) ucnt[s] = 0;
    if (AUXON) { mbar_init(abar, 1); *acnt = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int job = jobs ? jobs[blockIdx.x] : blockIdx.x;
  const int tx = job % G.ntx, ty = (job / G.ntx) % G.nty, sz = job / (G.ntx * G.nty);
  const int x0 = G.lox + tx * TXC, y0 = G.loy + ty * TYC, z0 = G.loz + sz * G.NZ;
  const int NZ = G.NZ;
  const size_t no = (size_t)G.n_owned;
  const long long pt0 = plane_term(G, z0);

  // one thread streams plane p of the tile into its stage: 8 bulk copies of 4 consecutive cells (TMA engine)
  // Both issue routines are called by a WHOLE warp: lane 0 arms the barrier, then one lane per bulk copy (the index
  // arithmetic of 8 - 24 copies on a single lane would hold that warp back by a good part of a plane period).
  auto issue_plane = [&](int p) {
    unsigned long long *bar = &ubar[p % NS];
    if (lane == 0) mbar_expect_tx(bar, (unsigned)(TC * ND * sizeof(T)));
    __syncwarp();
    const long long sh = plane_term(G, z0 + p) - pt0;
    T *stage = U + (size_t)(p % NS) * TC * ND;
    if (lane < TC / 4) {
      const int g = lane;
      const int ci = cell_index(G, x0 + slot_cx(4 * g), y0 + slot_cy(4 * g), z0);
      bulk_g2s(stage + (size_t)4 * g * ND, src + (size_t)(ci + sh) * ND, (unsigned)(4 * ND * sizeof(T)), bar);
    }
  };
  auto issue_aux = [&](int p) {
    const int nact = 1 + (E.d ? 1 : 0) + (E.xold ? 1 : 0);
    if (lane == 0) mbar_expect_tx(abar, (unsigned)(nact * TC * ND * sizeof(T)));
    __syncwarp();
    const long long sh = plane_term(G, z0 + p) - pt0;
    static_assert(3 * (TC / 4) <= 32, "one lane per bulk copy");
    if (lane < 3 * (TC / 4)) {
      const int a = lane / (TC / 4), g = lane - a * (TC / 4);
      const T *arr = a == 0 ? E.rhs : (a == 1 ? E.d : E.xold);
      if (arr) {
        const int ci = cell_index(G, x0 + slot_cx(4 * g), y0 + slot_cy(4 * g), z0);
        bulk_g2s(AUX + (size_t)a * TC * ND + (size_t)4 * g * ND, arr + (size_t)(ci + sh) * ND, (unsigned)(4 * ND * sizeof(T)), abar);
      }
    }
  };

  if (role == 1) {
    // ------------------------------------------------------------------ Z warps (column threads)
    const int zw = ridx;
    int coff[R::NCOL], cofp[R::NCOL], cxy[R::NCOL];   // offset of the column inside a u stage / padded plane; global element offset at z0
    bool cval[R::NCOL];
    int cij0 = 0;
#pragma unroll
    for (int m = 0; m < R::NCOL; ++m) {
      const int unit = zw + R::NZW * m;
      cval[m] = unit < NL;
      const int id = (cval[m] ? unit : 0) * 32 + lane;
      const int slot = id / NL, ij = id - slot * NL;
      coff[m] = slot * ND + ij;
      cofp[m] = slot * NDP + ij;
      cxy[m] = cell_index(G, x0 + slot_cx(slot), y0 + slot_cy(slot), z0) * ND + ij;   // element offset at plane z0
      (void)cij0;
    }
    T uc[R::NCOL][n], Gt[R::NCOL], Vt[R::NCOL];
    auto apply_mz_store = [&](int p) {   // dst(plane p) = Mz q
      const T *Ws = W + (size_t)(p & 1) * TC * NDP;
      const T *Upl = UP + (size_t)(p & 1) * TC * NDP;   // PAD: plane p of u is still there (this thread wrote these columns)
      const long long sh = (plane_term(G, z0 + p) - pt0) * ND;
      if (AUXON) mbar_wait(abar, (unsigned)(p & 1));
#pragma unroll
      for (int m = 0; m < R::NCOL; ++m) {
        if (!cval[m]) continue;
        T qv[n];
#pragma unroll
        for (int k = 0; k < n; ++k) qv[k] = Ws[cofp[m] + k * NL];
        const long long off = (long long)cxy[m] + sh;
        T *o = dst + off;
#pragma unroll
        for (int k = 0; k < n; ++k) {
          T acc = ADD ? o[k * NL] : T(0);
#pragma unroll
          for (int kk = 0; kk < n; ++kk) acc = fma(P.Mh[2][k * n + kk], qv[kk], acc);
          if (E.on) {   // fused Chebyshev / residual update: dst = a x + b xold + c d .* (rhs - A x)
            if (PAD) {
              T v = E.c * (AUX[coff[m] + k * NL] - acc);
              if (E.d) v *= AUX[TC * ND + coff[m] + k * NL];
              v += E.a * Upl[cofp[m] + k * NL];
              if (E.xold) v += E.b * AUX[2 * TC * ND + coff[m] + k * NL];
              acc = v;
            } else {
              T v = E.c * (E.rhs[off + k * NL] - acc);
              if (E.d) v *= E.d[off + k * NL];
              v += E.a * src[off + k * NL];
              if (E.xold) v += E.b * E.xold[off + k * NL];
              acc = v;
            }
          }
          o[k * NL] = acc;
        }
      }
      if (AUXON) {   // the last Z warp done with this plane's AUX starts the copy of the next one
        __syncwarp();
        int last = 0;
        if (lane == 0) {
          const int old = atomicAdd(acnt, 1);
          if (old == R::NZW - 1) { *acnt = 0; __threadfence_block(); last = 1; }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && p + 1 < NZ) issue_aux(p + 1);
      }
    };
    // one group of columns: finalise plane q (held in uc) with the bottom traces of the plane in stage Un
    auto z_group = [&](auto m0c, auto m1c, const T *Un, T *Ws, T *Ups) {
      constexpr int m0 = decltype(m0c)::value, m1 = decltype(m1c)::value, NG = m1 - m0;
      T nw[NG][n], gb[NG], gt[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int k = 0; k < n; ++k) nw[g][k] = cval[m0 + g] ? Un[coff[m0 + g] + k * NL] : T(0);
#pragma unroll
      for (int g = 0; g < NG; ++g) { gb[g] = 0; gt[g] = 0; }
#pragma unroll
      for (int k = 0; k < n; ++k)
#pragma unroll
        for (int g = 0; g < NG; ++g) { gb[g] = fma(P.g[0][k], nw[g][k], gb[g]); gt[g] = fma(P.g[1][k], uc[m0 + g][k], gt[g]); }
#pragma unroll
      for (int k = 0; k < n; ++k) {
        T acc[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) acc[g] = P.cG[2][0][k] * Gt[m0 + g];
#pragma unroll
        for (int g = 0; g < NG; ++g) acc[g] = fma(P.cV[2][0][k], Vt[m0 + g], acc[g]);
#pragma unroll
        for (int g = 0; g < NG; ++g) acc[g] = fma(P.cG[2][1][k], gb[g], acc[g]);
#pragma unroll
        for (int g = 0; g < NG; ++g) acc[g] = fma(P.cV[2][1][k], nw[g][0], acc[g]);
#pragma unroll
        for (int kk = 0; kk < n; ++kk)
#pragma unroll
          for (int g = 0; g < NG; ++g) acc[g] = fma(P.Kt[2][k * n + kk], uc[m0 + g][kk], acc[g]);   // kappa sits on the diagonal
#pragma unroll
        for (int g = 0; g < NG; ++g)
          if (cval[m0 + g]) {
            Ws[cofp[m0 + g] + k * NL] = acc[g];
            if (PAD) Ups[cofp[m0 + g] + k * NL] = uc[m0 + g][k];
          }
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        Gt[m0 + g] = gt[g]; Vt[m0 + g] = uc[m0 + g][n - 1];
#pragma unroll
        for (int k = 0; k < n; ++k) uc[m0 + g][k] = nw[g][k];
      }
    };
    // ghost plane below the segment: only its top traces are needed
#pragma unroll
    for (int m = 0; m < R::NCOL; ++m) {
      Gt[m] = 0; Vt[m] = 0;
      if (!cval[m]) continue;
      const int cell = cxy[m] / ND;
      const T *p = src + (size_t)nbr[4 * no + cell] * ND + (cxy[m] - cell * ND);
      T gsum = 0, last = 0;
#pragma unroll
      for (int k = 0; k < n; ++k) { const T x = __ldg(p + k * NL); gsum = fma(P.g[1][k], x, gsum); last = x; }
      Gt[m] = gsum; Vt[m] = last;
    }
    // PAD only: plane `pl` now lives in the registers of this warp; the last Z warp to say so refills its stage
    auto z_release_stage = [&](int pl) {
      if (!PAD) return;
#pragma unroll
      for (int m = 0; m < R::NCOL; ++m)
#pragma unroll
        for (int k = 0; k < n; ++k) keep(uc[m][k]);
      __syncwarp();
      int last = 0;
      if (lane == 0) {
        const int old = atomicAdd(&ucnt[pl % NS], 1);
        if (old == R::NZW - 1) { ucnt[pl % NS] = 0; __threadfence_block(); last = 1; }
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last && pl + NS < NZ) issue_plane(pl + NS);
    };
    mbar_wait(&ubar[0], 0);
#pragma unroll
    for (int m = 0; m < R::NCOL; ++m)
#pragma unroll
      for (int k = 0; k < n; ++k) uc[m][k] = cval[m] ? U[coff[m] + k * NL] : T(0);
    z_release_stage(0);
    for (int q = 0; q < NZ; ++q) {
      T *Un = U + (size_t)((q + 1) % NS) * TC * ND;
      T *Ws = W + (size_t)(q & 1) * TC * NDP;
      T *Ups = UP + (size_t)(q & 1) * TC * NDP;
      if (q + 1 < NZ) {
        mbar_wait(&ubar[(q + 1) % NS], (unsigned)(((q + 1) / NS) & 1));
      } else {
        // ghost plane above the segment: each thread copies its own columns into the (free) stage, then the
        // common path reads them back
        const long long shl = plane_term(G, z0 + NZ - 1) - pt0;
#pragma unroll
        for (int m = 0; m < R::NCOL; ++m) {
          if (!cval[m]) continue;
          const int cell = cxy[m] / ND;
          const T *p = src + (size_t)nbr[5 * no + (cell + shl)] * ND + (cxy[m] - cell * ND);
#pragma unroll
          for (int k = 0; k < n; ++k) Un[coff[m] + k * NL] = __ldg(p + k * NL);
        }
      }
      constexpr int MH = (R::NCOL + 1) / 2;
      z_group(std::integral_constant<int, 0>{}, std::integral_constant<int, MH>{}, Un, Ws, Ups);
      z_group(std::integral_constant<int, MH>{}, std::integral_constant<int, R::NCOL>{}, Un, Ws, Ups);
      if (q + 1 < NZ) z_release_stage(q + 1);
      bar_arrive(B_FULL + (q & 1), CNT_FULL);
      if (q >= 1) {
        bar_sync(B_QREADY + ((q - 1) & 1), CNT_QREADY);
        apply_mz_store(q - 1);
      }
    }
    bar_sync(B_QREADY + ((NZ - 1) & 1), CNT_QREADY);
    apply_mz_store(NZ - 1);
  } else if (role == 0) {
    // ------------------------------------------------------------------ XY warps (layer threads)
    // The n x n layer is streamed from shared memory row by row (y sweep) and column by column (x sweep) and
    // both sweeps are accumulated as outer products into one static register tile: the loops stay rolled, so
    // the instruction footprint of this role is a fraction of the unrolled form.
    const int k = ridx, c = lane, cx = slot_cx(c), cy = slot_cy(c);
    if (k == 0) {
      for (int p = 0; p < NS && p < NZ; ++p) issue_plane(p);
      if (AUXON) issue_aux(0);
    }
    // neighbour descriptors per (direction, side): where G' and V' of the facing cell live
    //   inside the tile: G' in TX/TY (stride TC per face point), V' in the u stage (stride n or 1)
    //   halo:            (V', G') pairs in the halo buffer (stride 2)
    int g_off[2][2], v_off[2][2];
    bool ins[2][2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      ins[0][s] = s ? (cx < TXC - 1) : (cx > 0);
      ins[1][s] = s ? (cy < TYC - 1) : (cy > 0);
      const int cnx = slot_of(s ? cx + 1 : cx - 1, cy), cny = slot_of(cx, s ? cy + 1 : cy - 1);
      // x faces: face point = row j (of layer k)
      g_off[0][s] = ins[0][s] ? (int)(TX - U) + ((1 - s) * NL + n * k) * TC + cnx : (int)(HXY - U) + ((s * TYC + cy) * NL + n * k) * 2 + 1;
      v_off[0][s] = ins[0][s] ? cnx * NDP + k * NL + (1 - s) * (n - 1) : (int)(HXY - U) + ((s * TYC + cy) * NL + n * k) * 2;
      // y faces: face point = column i
      g_off[1][s] = ins[1][s] ? (int)(TY - U) + ((1 - s) * NL + n * k) * TC + cny : (int)(HXY - U) + 2 * TYC * NL * 2 + ((s * TXC + cx) * NL + n * k) * 2 + 1;
      v_off[1][s] = ins[1][s] ? cny * NDP + k * NL + (1 - s) * (n - 1) * n : (int)(HXY - U) + 2 * TYC * NL * 2 + ((s * TXC + cx) * NL + n * k) * 2;
    }
    for (int p = 0; p < NZ; ++p) {
      const int st = p & 1;
      bar_sync(B_FULL + st, CNT_FULL);
      if (!PAD) mbar_wait(&ubar[p % NS], (unsigned)((p / NS) & 1));   // already complete: orders the async-proxy writes
      const int ustage = PAD ? (int)(UP - U) + st * TC * NDP : (p % NS) * TC * ND;   // this plane's u, relative to U
      T *Wp = W + (size_t)st * TC * NDP + c * NDP + k * NL;
      const T *Up = U + ustage + c * NDP + k * NL;
      T acc[n][n], gx[2][n], gy[2][n];
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int i = 0; i < n; ++i) acc[j][i] = Wp[j * n + i];
#pragma unroll
      for (int a = 0; a < n; ++a) { gx[0][a] = 0; gx[1][a] = 0; gy[0][a] = 0; gy[1][a] = 0; }
      // y sweep: rows of u, acc[jj][i] += Kt_y[jj][j] u[j][i]; traces of the y faces
      if (XY_UNROLLED) {
        // the layer sits in registers; both sweeps and the four face traces straight-line (more ILP, larger code)
        T u[n][n];
#pragma unroll
        for (int j = 0; j < n; ++j)
#pragma unroll
          for (int i = 0; i < n; ++i) u[j][i] = Up[j * n + i];
#pragma unroll
        for (int j = 0; j < n; ++j)
#pragma unroll
          for (int i = 0; i < n; ++i) {
            T a = acc[j][i];
#pragma unroll
            for (int l = 0; l < n; ++l) a = fma(P.Kt[0][i * n + l], u[j][l], a);
#pragma unroll
            for (int l = 0; l < n; ++l) a = fma(P.Kt[1][j * n + l], u[l][i], a);
            acc[j][i] = a;
          }
#pragma unroll
        for (int a = 0; a < n; ++a)
#pragma unroll
          for (int l = 0; l < n; ++l) {
            gx[0][a] = fma(P.g[0][l], u[a][l], gx[0][a]); gx[1][a] = fma(P.g[1][l], u[a][l], gx[1][a]);
            gy[0][a] = fma(P.g[0][l], u[l][a], gy[0][a]); gy[1][a] = fma(P.g[1][l], u[l][a], gy[1][a]);
          }
      } else {
      T nxt[n];
#pragma unroll
      for (int i = 0; i < n; ++i) nxt[i] = Up[i];
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        T row[n], kc[n];
#pragma unroll
        for (int i = 0; i < n; ++i) row[i] = nxt[i];
        const int jn = j + 1 < n ? j + 1 : 0;          // software pipelining: the next row is in flight during the FMAs
#pragma unroll
        for (int i = 0; i < n; ++i) nxt[i] = Up[jn * n + i];
#pragma unroll
        for (int jj = 0; jj < n; ++jj) kc[jj] = P.Kt[1][jj * n + j];
        const T g0 = P.g[0][j], g1 = P.g[1][j];
#pragma unroll
        for (int jj = 0; jj < n; ++jj)
#pragma unroll
          for (int i = 0; i < n; ++i) acc[jj][i] = fma(kc[jj], row[i], acc[jj][i]);
#pragma unroll
        for (int i = 0; i < n; ++i) { gy[0][i] = fma(g0, row[i], gy[0][i]); gy[1][i] = fma(g1, row[i], gy[1][i]); }
      }
      // x sweep: columns of u, acc[j][ii] += Kt_x[ii][i] u[j][i]; traces of the x faces
#pragma unroll
      for (int j = 0; j < n; ++j) nxt[j] = Up[j * n];
#pragma unroll 1
      for (int i = 0; i < n; ++i) {
        T col[n], kc[n];
#pragma unroll
        for (int j = 0; j < n; ++j) col[j] = nxt[j];
        const int in_ = i + 1 < n ? i + 1 : 0;
#pragma unroll
        for (int j = 0; j < n; ++j) nxt[j] = Up[j * n + in_];
#pragma unroll
        for (int ii = 0; ii < n; ++ii) kc[ii] = P.Kt[0][ii * n + i];
        const T g0 = P.g[0][i], g1 = P.g[1][i];
#pragma unroll
        for (int j = 0; j < n; ++j)
#pragma unroll
          for (int ii = 0; ii < n; ++ii) acc[j][ii] = fma(kc[ii], col[j], acc[j][ii]);
#pragma unroll
        for (int j = 0; j < n; ++j) { gx[0][j] = fma(g0, col[j], gx[0][j]); gx[1][j] = fma(g1, col[j], gx[1][j]); }
      }
      }
      if (p > 0) bar_sync(B_XY, CNT_XY);   // every layer thread has read the previous plane's traces: TX / TY may be overwritten
#pragma unroll
      for (int a = 0; a < n; ++a) {
        TX[(0 * NL + a + n * k) * TC + c] = gx[0][a];
        TX[(1 * NL + a + n * k) * TC + c] = gx[1][a];
        TY[(0 * NL + a + n * k) * TC + c] = gy[0][a];
        TY[(1 * NL + a + n * k) * TC + c] = gy[1][a];
      }
      bar_sync(B_XY, CNT_XY);
      // neighbour terms; stage / buffer dependent part of the descriptors
      const int hsh = st * HBUF;
#pragma unroll XY_UNROLLED ? 2 : 1
      for (int s = 0; s < 2; ++s) {
        const bool in = s ? ins[0][1] : ins[0][0];
        const T *gp = U + (s ? g_off[0][1] : g_off[0][0]) + (in ? 0 : hsh);
        const T *vp = U + (s ? v_off[0][1] : v_off[0][0]) + (in ? ustage : hsh);
        const int gs = in ? TC : 2, vs = in ? n : 2;
        T cg[n], cv[n];
#pragma unroll
        for (int i = 0; i < n; ++i) { cg[i] = P.cG[0][s][i]; cv[i] = P.cV[0][s][i]; }
#pragma unroll
        for (int j = 0; j < n; ++j) {
          const T gn = gp[j * gs], vn = vp[j * vs];
#pragma unroll
          for (int i = 0; i < n; ++i) acc[j][i] = fma(cv[i], vn, fma(cg[i], gn, acc[j][i]));
        }
      }
#pragma unroll XY_UNROLLED ? 2 : 1
      for (int s = 0; s < 2; ++s) {
        const bool in = s ? ins[1][1] : ins[1][0];
        const T *gp = U + (s ? g_off[1][1] : g_off[1][0]) + (in ? 0 : hsh);
        const T *vp = U + (s ? v_off[1][1] : v_off[1][0]) + (in ? ustage : hsh);
        const int gs = in ? TC : 2, vs = in ? 1 : 2;
        T cg[n], cv[n];
#pragma unroll
        for (int j = 0; j < n; ++j) { cg[j] = P.cG[1][s][j]; cv[j] = P.cV[1][s][j]; }
#pragma unroll
        for (int i = 0; i < n; ++i) {
          const T gn = gp[i * gs], vn = vp[i * vs];
#pragma unroll
          for (int j = 0; j < n; ++j) acc[j][i] = fma(cv[j], vn, fma(cg[j], gn, acc[j][i]));
        }
      }
      // all shared-memory reads of this plane's u and traces are done: make sure they have landed
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int i = 0; i < n; ++i) keep(acc[j][i]);
      if (!PAD) {   // the last XY warp to finish reading plane p refills its stage with plane p + NS
        int last = 0;
        if (lane == 0) {
          const int old = atomicAdd(&ucnt[p % NS], 1);
          if (old == R::NXY - 1) { ucnt[p % NS] = 0; last = 1; }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && p + NS < NZ) issue_plane(p + NS);
      }
      // mass in x, then y
      T tmp[n][n];
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int i = 0; i < n; ++i) {
          T a = P.Mh[0][i * n] * acc[j][0];
#pragma unroll
          for (int l = 1; l < n; ++l) a = fma(P.Mh[0][i * n + l], acc[j][l], a);
          tmp[j][i] = a;
        }
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int i = 0; i < n; ++i) {
          T a = P.Mh[1][j * n] * tmp[0][i];
#pragma unroll
          for (int l = 1; l < n; ++l) a = fma(P.Mh[1][j * n + l], tmp[l][i], a);
          Wp[j * n + i] = a;
        }
      bar_arrive(B_QREADY + st, CNT_QREADY);
    }
  } else {
    // ------------------------------------------------------------------ H warps (halo traces)
    // The (halo cell, face point) items are dealt to the 64 threads in slots whose face (direction, side) is a
    // compile-time constant, so that every slot is a short straight-line sequence: one neighbour-table lookup
    // (prefetched one plane ahead), n loads, n FMAs, one (V', G') store.
    constexpr int NT = R::NH * 32;
    constexpr int SX = (TYC * NL + NT - 1) / NT, SY = (TXC * NL + NT - 1) / NT;   // slots per x side / y side
    constexpr int NSLOT = 2 * SX + 2 * SY;
    const int t = ridx * 32 + lane;
    int hcell[NSLOT], hline[NSLOT], hcn[NSLOT];   // adjacent tile cell (index at plane z0), line offset, halo cell
    auto for_slots = [&](auto &&f) {
      // f(slot, direction, side, item-within-side)
#define DGX_SLOT(SL) f(std::integral_constant<int, SL>{})
      [&]<int... I>(std::integer_sequence<int, I...>) { (DGX_SLOT(I), ...); }(std::make_integer_sequence<int, NSLOT>{});
#undef DGX_SLOT
    };
    auto slot_dir = [](int sl) { return sl < 2 * SX ? 0 : 1; };
    auto slot_side = [](int sl) { return sl < 2 * SX ? sl / SX : (sl - 2 * SX) / SY; };
    auto slot_round = [](int sl) { return sl < 2 * SX ? sl % SX : (sl - 2 * SX) % SY; };
    for_slots([&](auto slc) {
      constexpr int sl = decltype(slc)::value;
      constexpr int dir = sl < 2 * SX ? 0 : 1, side = dir == 0 ? sl / SX : (sl - 2 * SX) / SY, rnd = dir == 0 ? sl % SX : (sl - 2 * SX) % SY;
      constexpr int per_side = (dir == 0 ? TYC : TXC) * NL;
      int e = t + rnd * NT;                    // item within this side; clamp: invalid items redo the last one
      if (e >= per_side) e = per_side - 1;
      const int cellrow = e / NL, fp = e - cellrow * NL;
      const int hcx = dir == 0 ? (side ? TXC - 1 : 0) : cellrow, hcy = dir == 0 ? cellrow : (side ? TYC - 1 : 0);
      hcell[sl] = cell_index(G, x0 + hcx, y0 + hcy, z0);
      hline[sl] = dir == 0 ? fp * n : (fp % n) + (fp / n) * NL;
      hcn[sl] = nbr[(size_t)(2 * dir + side) * no + hcell[sl]];
    });
    (void)slot_dir; (void)slot_side; (void)slot_round;
    for (int p = 0; p < NZ + 2; ++p) {
      if (p >= 2) bar_sync(B_QREADY + (p & 1), CNT_QREADY);   // plane p - 2 is done: buffer p & 1 is free
      if (p < NZ) {
        T x[NSLOT][n];
        for_slots([&](auto slc) {
          constexpr int sl = decltype(slc)::value;
          constexpr int str = sl < 2 * SX ? 1 : n;
          const T *ptr = src + (size_t)hcn[sl] * ND + hline[sl];
#pragma unroll
          for (int l = 0; l < n; ++l) x[sl][l] = __ldg(ptr + l * str);
        });
        const long long sh = plane_term(G, z0 + (p + 1 < NZ ? p + 1 : p)) - pt0;
        for_slots([&](auto slc) {
          constexpr int sl = decltype(slc)::value;
          constexpr int dir = sl < 2 * SX ? 0 : 1, side = dir == 0 ? sl / SX : (sl - 2 * SX) / SY;
          hcn[sl] = nbr[(size_t)(2 * dir + side) * no + (hcell[sl] + sh)];
        });
        T *hb = HXY + (size_t)(p & 1) * HBUF;
        for_slots([&](auto slc) {
          constexpr int sl = decltype(slc)::value;
          constexpr int dir = sl < 2 * SX ? 0 : 1, side = dir == 0 ? sl / SX : (sl - 2 * SX) / SY, rnd = dir == 0 ? sl % SX : (sl - 2 * SX) % SY;
          constexpr int per_side = (dir == 0 ? TYC : TXC) * NL;
          constexpr int base = (dir == 0 ? 0 : 2 * TYC * NL) + side * per_side;
          T gsum = 0;
#pragma unroll
          for (int l = 0; l < n; ++l) gsum = fma(P.g[1 - side][l], x[sl][l], gsum);    // the neighbour's facing side is 1 - side
          const int e = t + rnd * NT;
          if (e < per_side) {
            hb[2 * (base + e)] = side ? x[sl][0] : x[sl][n - 1];
            hb[2 * (base + e) + 1] = gsum;
          }
        });
        bar_arrive(B_FULL + (p & 1), CNT_FULL);
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
    if (d == 2) for (int i = 0; i < n; ++i) Kt[i * n + i] += pressure_kappa(op);   // the Helmholtz term rides on the z sweep
    for (int i = 0; i < n * n; ++i) { F.Kt[d][i] = (T)Kt[i]; F.Mh[d][i] = (T)(h * s.M[i]); }
  }
  for (int sd = 0; sd < 2; ++sd)
    for (int j = 0; j < n; ++j) F.g[sd][j] = (T)((sd ? 1.0 : -1.0) * s.g[sd][j]);
  F.kappa = (T)pressure_kappa(op);
}

#include "pressure_march_v2.cuh"
#include "pressure_march_v3.cuh"

// which marching kernel runs the plain (no add, no fused update) Q4 fp64 operator: DGX_MARCH = v1 | v2 | v2eo | v2r | v2eor | v2eol | v3 (default) | v3x5 | v3x6 | v3b2 | v3b32
inline int march_generation() {   // read on every call (tests and the bench switch it inside one process)
  const char *e = getenv("DGX_MARCH");
  if (!e) return 10;
  if (!strcmp(e, "v1")) return 0;      // round-1 kernel (named barriers, register tiles, STG)
  if (!strcmp(e, "v2")) return 1;      // mbarrier hand-offs, shuffled traces, bulk stores; plain contractions, layer tile in registers
  if (!strcmp(e, "v2eo")) return 2;    // + even-odd contractions
  if (!strcmp(e, "v2r")) return 3;     // plain contractions, streaming XY role
  if (!strcmp(e, "v2eol")) return 5;   // even-odd, XY role in lockstep batches
  if (!strcmp(e, "v3")) return 10;     // v3: single Z visit (Mz folded), XY tasks on 7 warps, halo traces by the Z warps, batches of 3 lines
  if (!strcmp(e, "v3x5")) return 11;   // 5 XY + 5 Z warps
  if (!strcmp(e, "v3x6")) return 12;   // 6 XY + 5 Z warps
  if (!strcmp(e, "v3b2")) return 13;   // as v3, batches of 2 lines
  if (!strcmp(e, "v3b32")) return 14;  // as v3, rows in batches of 3, columns in batches of 2
  if (!strcmp(e, "v3zf")) return 16;   // as v3, Z warps on the low warp ids
  if (!strcmp(e, "v2eor")) return 4;   // even-odd, streaming XY role with a 3-stage software pipeline (round-2 default before v3)
  return 10;                           // v3 (default)
}

template <int n, typename T, bool EO, int XYV>
int launch_march2(dgx_op *op, dgx_vec *dst, const dgx_vec *src, const FastGeom &G_, const int *d_jobs, int n_jobs, cudaStream_t stream) {
  using R = Roles2<n>;
  constexpr int NL = n * n, ND = NL * n;
  FastGeom G = G_;
  { const char *e = getenv("DGX_MARCH_WAIT"); G.wait_mode = e ? atoi(e) : 0; }
  FastParams<T, n> F;
  fill_fast_params<T, n>(F, op);
  EOParams<T, n> Q;
  {
    FastParams<double, n> Fd;
    fill_fast_params<double, n>(Fd, op);
    if (EO && eo_asymmetry<n>(Fd) > 1e-13) { set_error("marching kernel: 1-D tables are not centro-symmetric"); return DGX_ERR_INVALID; }
    fill_eo_params<T, n>(Q, Fd);
  }
  constexpr size_t smem = sizeof(T) * (size_t)((R::NSTAGE + 3) * TC * ND + 2 * (2 * TYC + 2 * TXC) * NL * 2) + 8 * (R::NSTAGE + 6) + 4 * R::NSTAGE + 16;
  int dev = 0;
  DGX_CUDA_CHECK(cudaGetDevice(&dev));
  static bool configured[64] = {};
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    DGX_CUDA_CHECK(cudaFuncSetAttribute(pressure_march2_kernel<n, T, EO, XYV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev] = true;
  }
  const int grid = d_jobs ? n_jobs : G.ntx * G.nty * G.nsz;
  if (grid == 0) return DGX_OK;
  pressure_march2_kernel<n, T, EO, XYV><<<grid, R::NW * 32, smem, stream>>>(F, Q, G, op->mf->d_nbr, (const T *)src->d, (T *)dst->d, d_jobs);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}


template <int n, typename T, int NXY, int NBX, int NBY, bool ZFIRST = false, int MINB = 1>
int launch_march3(dgx_op *op, dgx_vec *dst, const dgx_vec *src, const FastGeom &G, const int *d_jobs, int n_jobs, cudaStream_t stream,
                  const GhostSrc *gsrc = nullptr) {
  using R = Roles3<n, NXY>;
  constexpr int NL = n * n, ND = NL * n;
  EOParams<T, n> Q;
  ZTop<T, n> ZT;
  {
    FastParams<double, n> Fd;
    fill_fast_params<double, n>(Fd, op);
    if (eo_asymmetry<n>(Fd) > 1e-13) { set_error("marching kernel: 1-D tables are not centro-symmetric"); return DGX_ERR_INVALID; }
    fill_eo_params_v3<T, n>(Q, ZT, Fd);
  }
  constexpr size_t smem = sizeof(T) * (size_t)((R::NSU + R::NSW) * TC * ND + R::NSH * (2 * TYC + 2 * TXC) * NL * 2) + 8 * (4 * R::NSU + 4) + 4 * R::NSU + 16;
  int dev = 0;
  DGX_CUDA_CHECK(cudaGetDevice(&dev));
  static bool configured[64] = {};
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    DGX_CUDA_CHECK(cudaFuncSetAttribute(pressure_march3_kernel<n, T, NXY, NBX, NBY, ZFIRST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev] = true;
  }
  const int grid = d_jobs ? n_jobs : G.ntx * G.nty * G.nsz;
  if (grid == 0) return DGX_OK;
  pressure_march3_kernel<n, T, NXY, NBX, NBY, ZFIRST, MINB><<<grid, R::NW * 32, smem, stream>>>(Q, ZT, G, op->mf->d_nbr, (const T *)src->d, (T *)dst->d, d_jobs, gsrc ? *gsrc : GhostSrc{});
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

template <int n, typename T>
int launch_march(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, const FastGeom &G, const int *d_jobs, int n_jobs, cudaStream_t stream,
                 const FusedUpdate *fu) {
  if constexpr (n == 5 && std::is_same<T, double>::value) {
    if (!add && !fu && march_generation() > 0)
      switch (march_generation()) {
        case 10: return launch_march3<n, T, 7, 3, 3>(op, dst, src, G, d_jobs, n_jobs, stream);
        case 11: return launch_march3<n, T, 5, 3, 3>(op, dst, src, G, d_jobs, n_jobs, stream);
        case 12: return launch_march3<n, T, 6, 3, 3>(op, dst, src, G, d_jobs, n_jobs, stream);
        case 13: return launch_march3<n, T, 7, 2, 2>(op, dst, src, G, d_jobs, n_jobs, stream);
        case 14: return launch_march3<n, T, 7, 3, 2>(op, dst, src, G, d_jobs, n_jobs, stream);
        case 16: return launch_march3<n, T, 7, 3, 3, true>(op, dst, src, G, d_jobs, n_jobs, stream);
        case 1: return launch_march2<n, T, false, 0>(op, dst, src, G, d_jobs, n_jobs, stream);
        case 3: return launch_march2<n, T, false, 1>(op, dst, src, G, d_jobs, n_jobs, stream);
        case 4: return launch_march2<n, T, true, 1>(op, dst, src, G, d_jobs, n_jobs, stream);
        case 5: return launch_march2<n, T, true, 2>(op, dst, src, G, d_jobs, n_jobs, stream);
        case 2: return launch_march2<n, T, true, 0>(op, dst, src, G, d_jobs, n_jobs, stream);
        default: return launch_march3<n, T, 7, 3, 3>(op, dst, src, G, d_jobs, n_jobs, stream);
      }
  }
  if constexpr (n == 3) {
    // Q2: the v3 design with 3 Z + 6 XY warps and 2 (fp64) / 3 (fp32) CTAs per SM: 134 -> 164 GDoF/s fp64, 155 -> 243 GDoF/s fp32 at
    // 256^3 cells (DGX_MARCH3_Q2=0: round-1 kernel)
    static const int q2 = [] { const char *e = getenv("DGX_MARCH3_Q2"); return e ? atoi(e) : 1; }();
    if (!add && !fu && q2) {
      if constexpr (std::is_same<T, double>::value) return launch_march3<n, T, 6, 3, 3, false, 2>(op, dst, src, G, d_jobs, n_jobs, stream);
      else return launch_march3<n, T, 6, 3, 3, false, 3>(op, dst, src, G, d_jobs, n_jobs, stream);
    }
  }
  if constexpr (n == 5 && std::is_same<T, float>::value) {
    // Q4 fp32: the v3 kernel as for fp64 (one CTA per SM): 212 -> 282 GDoF/s at 128^3 cells (DGX_MARCH3_F32=0: round-1 kernel)
    static const int f32 = [] { const char *e = getenv("DGX_MARCH3_F32"); return e ? atoi(e) : 1; }();
    if (!add && !fu && f32) return launch_march3<n, T, 7, 3, 3>(op, dst, src, G, d_jobs, n_jobs, stream);
  }
  using R = Roles<n>;
  constexpr int NL = n * n, ND = NL * n;
  FastParams<T, n> F;
  fill_fast_params<T, n>(F, op);
  constexpr int NDP = (n % 2 == 0) ? ND + 1 : ND;
  const size_t smem_plain = sizeof(T) * (size_t)(R::NSTAGE * TC * ND + (n % 2 == 0 ? 4 : 2) * TC * NDP + 4 * NL * TC + 2 * 2 * TYC * NL * 2 + 2 * 2 * TXC * NL * 2) + 8 * R::NSTAGE + 8 * ((4 * R::NSTAGE + 7) / 8) + 16;
  const size_t smem_fused = smem_plain + (n % 2 == 0 ? sizeof(T) * (size_t)(3 * TC * ND) : 0);   // + AUX planes of the fused update
  const size_t smem = fu ? smem_fused : smem_plain;
  int dev = 0;
  DGX_CUDA_CHECK(cudaGetDevice(&dev));
  static bool configured[64] = {};   // per device: the attribute belongs to the device's copy of the function
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    DGX_CUDA_CHECK(cudaFuncSetAttribute(pressure_march_kernel<n, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fused));
    DGX_CUDA_CHECK(cudaFuncSetAttribute(pressure_march_kernel<n, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_plain));
    configured[dev] = true;
  }
  const int grid = d_jobs ? n_jobs : G.ntx * G.nty * G.nsz;
  if (grid == 0) return DGX_OK;
  Epilogue<T> E;
  if (fu) {
    E.on = 1;
    E.xold = fu->xold ? (const T *)fu->xold->d : nullptr;
    E.d = fu->d ? (const T *)fu->d->d : nullptr;
    E.rhs = (const T *)fu->rhs->d;
    E.a = (T)fu->a; E.b = (T)fu->b; E.c = (T)fu->c;
  }
  if (add) pressure_march_kernel<n, T, true><<<grid, R::NW * 32, smem, stream>>>(F, G, op->mf->d_nbr, (const T *)src->d, (T *)dst->d, d_jobs, E);
  else pressure_march_kernel<n, T, false><<<grid, R::NW * 32, smem, stream>>>(F, G, op->mf->d_nbr, (const T *)src->d, (T *)dst->d, d_jobs, E);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

int march_dispatch(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, const FastGeom &G, const int *d_jobs, int n_jobs, cudaStream_t stream = nullptr,
                   const FusedUpdate *fu = nullptr) {
  dgx_mf *mf = op->mf;
  if (!stream) stream = mf->ctx->stream;
  if (mf->dtype == DGX_F64) {
    if (op->degree == 2) return launch_march<3, double>(op, dst, src, add, G, d_jobs, n_jobs, stream, fu);
    return op->degree == 4 ? launch_march<5, double>(op, dst, src, add, G, d_jobs, n_jobs, stream, fu) : launch_march<4, double>(op, dst, src, add, G, d_jobs, n_jobs, stream, fu);
  }
  if (op->degree == 2) return launch_march<3, float>(op, dst, src, add, G, d_jobs, n_jobs, stream, fu);
  return op->degree == 4 ? launch_march<5, float>(op, dst, src, add, G, d_jobs, n_jobs, stream, fu) : launch_march<4, float>(op, dst, src, add, G, d_jobs, n_jobs, stream, fu);
}

// geometry of the march for this operator, or false when the marching kernel does not apply
bool march_geometry(const dgx_op *op, FastGeom &G) {
  const dgx_mf *mf = op->mf;
  if (mf->dim != 3 || !mf->all_interior || !mf->box_ok || mf->general) return false;
  if (op->degree < 2 || op->degree > 4) return false;   // Q2: 84 -> 125 GDoF/s fp64, 128 -> 145 fp32 against the generic kernel (128^3)
  const Mesh &m = mf->mesh->m;
  G.L = mf->level;
  G.c0x = m.cells0[0]; G.c0y = m.cells0[1];
  G.lox = mf->box_lo[0]; G.loy = mf->box_lo[1]; G.loz = mf->box_lo[2];
  const int EX = mf->box_ext[0], EY = mf->box_ext[1], EZ = mf->box_ext[2];
  if (G.L < 3 || EX % TXC || EY % TYC || EZ < 2 || mf->first_cell % 4) return false;
  G.ntx = EX / TXC; G.nty = EY / TYC;
  // planes per job: as long as possible (each job pays ~2 extra planes of prologue / epilogue) but at least ~4 waves of jobs;
  // levels too small for that stay with the generic kernel, which parallelises over all cells
  int NZ = EZ;
  while (NZ > 64 && NZ % 2 == 0) NZ /= 2;
  while (NZ > 8 && NZ % 2 == 0 && (long long)G.ntx * G.nty * (EZ / NZ) < 4 * 148) NZ /= 2;
  if ((long long)G.ntx * G.nty * (EZ / NZ) < 2 * 148 && op->variant != 2) return false;
  G.NZ = NZ; G.nsz = EZ / NZ;
  G.first_cell = mf->first_cell;
  G.n_owned = (int)mf->n_owned;
  G.wait_mode = 0;
  return true;
}

}  // namespace

// fused operator + vector update through the marching kernel (its Z warps apply the update when they store the result)
int pressure_apply_fast_fused(dgx_op *op, dgx_vec *dst, const dgx_vec *x, const FusedUpdate &fu) {
  FastGeom G;
  if (op->variant == 1 || !march_geometry(op, G)) return DGX_ERR_UNSUPPORTED;
  return march_dispatch(op, dst, x, false, G, nullptr, 0, nullptr, &fu);
}


// Multi-rank vmult with the ghost import in flight (v3 kernel, Q4 fp64): the boundary cells are packed, the copy engines carry them to
// the neighbours over NVLink, and the marching kernel starts at once -- tile columns that read ghost cells from their first plane on
// (rank-boundary faces in x / y, the ghost plane below) are scheduled last, the others touch a ghost plane only at the end of their
// march; each warp waits for the neighbours' sequence numbers where it first needs a ghost cell and reads it from the mailbox.
int pressure_apply_fast_async(dgx_op *op, dgx_vec *dst, dgx_vec *src, bool add) {
  FastGeom G;
  dgx_mf *mf = op->mf;
  if (add || op->variant == 1 || mf->dtype != DGX_F64 || op->degree != 4 || march_generation() != 10 || !march_geometry(op, G)) return DGX_ERR_UNSUPPORTED;
  GhostSrc gs;
  int done = 0;
  int rc = p2p_ghost_import_async(src, &gs, &done);
  if (rc) return rc;
  if (!done) return DGX_ERR_UNSUPPORTED;
  if (!mf->d_jobs_async || mf->jobs_async_NZ != G.NZ) {
    const Mesh &m = mf->mesh->m;
    bool ghost_face[6];
    for (int f = 0; f < 6; ++f) {
      int c[3] = {mf->box_lo[0], mf->box_lo[1], mf->box_lo[2]};
      if (f & 1) c[f / 2] += mf->box_ext[f / 2] - 1;
      const long long local = m.index_of(mf->level, c) - mf->first_cell;
      ghost_face[f] = mf->nbr_host[(size_t)f * mf->n_owned + local] >= mf->n_owned;
    }
    std::vector<int> order[3];   // no ghost at all / ghost plane above only (read at the end) / ghosts from the first plane on
    for (int sz = 0; sz < G.nsz; ++sz)
      for (int ty = 0; ty < G.nty; ++ty)
        for (int tx = 0; tx < G.ntx; ++tx) {
          const bool early = (tx == 0 && ghost_face[0]) || (tx == G.ntx - 1 && ghost_face[1]) || (ty == 0 && ghost_face[2]) ||
                             (ty == G.nty - 1 && ghost_face[3]) || (sz == 0 && ghost_face[4]);
          const bool late = sz == G.nsz - 1 && ghost_face[5];
          order[early ? 2 : (late ? 1 : 0)].push_back(tx + G.ntx * (ty + G.nty * sz));
        }
    std::vector<int> all;
    for (auto &o : order) all.insert(all.end(), o.begin(), o.end());
    if (mf->d_jobs_async) cudaFree(mf->d_jobs_async);
    DGX_CUDA_CHECK(cudaMalloc(&mf->d_jobs_async, all.size() * sizeof(int)));
    DGX_CUDA_CHECK(cudaMemcpy(mf->d_jobs_async, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice));
    mf->jobs_async_NZ = G.NZ;
  }
  return launch_march3<5, double, 7, 3, 3>(op, dst, src, G, mf->d_jobs_async, G.ntx * G.nty * G.nsz, mf->ctx->stream, &gs);
}

// The marching kernel applies when every face is interior (fully periodic mesh) and the owned cells form a
// box of the hierarchical order whose extents are multiples of the tile.
int pressure_apply_fast(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add) {
  FastGeom G;
  if (!march_geometry(op, G)) return DGX_ERR_UNSUPPORTED;
  return march_dispatch(op, dst, src, add, G, nullptr, 0);
}

// Multi-rank: MatrixFree::loop's update_ghost_values_start / cells without ghost dependence / finish / remaining cells
// (SURVEY 9.10).  Tile columns on the surface of the owned box may read ghost cells (halo traces, ghost planes); all the
// others run while the ghost import is in flight on the communication stream.
int pressure_apply_fast_overlapped(dgx_op *op, dgx_vec *dst, dgx_vec *src, bool add) {
  FastGeom G;
  if (!march_geometry(op, G)) return DGX_ERR_UNSUPPORTED;
  dgx_mf *mf = op->mf;
  if (!mf->d_jobs || mf->jobs_NZ != G.NZ) {
    // which faces of the owned box border another rank (the partition is a box, so one cell per face tells)
    const Mesh &m = mf->mesh->m;
    bool ghost_face[6];
    for (int f = 0; f < 6; ++f) {
      int c[3] = {mf->box_lo[0], mf->box_lo[1], mf->box_lo[2]};
      if (f & 1) c[f / 2] += mf->box_ext[f / 2] - 1;
      const long long local = m.index_of(mf->level, c) - mf->first_cell;
      ghost_face[f] = mf->nbr_host[(size_t)f * mf->n_owned + local] >= mf->n_owned;
    }
    std::vector<int> interior, boundary;
    for (int sz = 0; sz < G.nsz; ++sz)
      for (int ty = 0; ty < G.nty; ++ty)
        for (int tx = 0; tx < G.ntx; ++tx) {
          const bool surf = (tx == 0 && ghost_face[0]) || (tx == G.ntx - 1 && ghost_face[1]) || (ty == 0 && ghost_face[2]) ||
                            (ty == G.nty - 1 && ghost_face[3]) || (sz == 0 && ghost_face[4]) || (sz == G.nsz - 1 && ghost_face[5]);
          (surf ? boundary : interior).push_back(tx + G.ntx * (ty + G.nty * sz));
        }
    mf->n_jobs_interior = (int)interior.size(); mf->n_jobs_boundary = (int)boundary.size();
    interior.insert(interior.end(), boundary.begin(), boundary.end());
    if (mf->d_jobs) cudaFree(mf->d_jobs);
    DGX_CUDA_CHECK(cudaMalloc(&mf->d_jobs, interior.size() * sizeof(int)));
    DGX_CUDA_CHECK(cudaMemcpy(mf->d_jobs, interior.data(), interior.size() * sizeof(int), cudaMemcpyHostToDevice));
    mf->jobs_NZ = G.NZ;
  }
  // communication stream: pack, NCCL send/recv, then the tile columns that read ghosts; compute stream: all other tile
  // columns.  The two kernels write disjoint cells of dst and fill each other's idle SMs; the compute stream joins at the end.
  dgx_ctx *ctx = mf->ctx;
  int rc = ghost_exchange_start(src);
  if (rc) return rc;
  rc = march_dispatch(op, dst, src, add, G, mf->d_jobs + mf->n_jobs_interior, mf->n_jobs_boundary, ctx->comm_stream);
  if (rc) return rc;
  DGX_CUDA_CHECK(cudaEventRecord(ctx->ev_ghosts_done, ctx->comm_stream));
  rc = march_dispatch(op, dst, src, add, G, mf->d_jobs, mf->n_jobs_interior);
  if (rc) return rc;
  return ghost_exchange_wait(ctx);
}

}  // namespace dgx
