// Third-generation marching kernel of the pressure operator (odd n; instantiated for Q4 fp64 = the headline workload).
// Included by pressure_fast.cu inside its anonymous namespace, after pressure_march_v2.cuh (uses its line operators,
// EOParams, the mbarrier / bulk-copy helpers and FastGeom).
//
// What changed against v2 (ncu of v2: l1tex 58 %, FP64 pipe 40 %, issue slots 40 %; XY warps the critical path at
// 0.19 IPC; Z warps visit every plane twice and wait 22 % of their time for the XY warps):
//
//   * one visit of the Z role per plane.  The mass matrix of the z direction commutes with the x / y line operators
//     (uniform Cartesian cells), so with u^ = Mz u
//         A u = Mx My [ (Mz Lz) u + (Lx + Ly) u^ ],      Lz u = Kt'_z u + face terms (kappa folded in)
//     and (Mz Lz) =: Az is ONE folded n x (n + 4) line operator.  The Z warps (lane = cell, warp = row j, 5 columns per
//     thread) write u^ over the u plane in place and Az u into W; the XY role then finishes the plane alone: x / y line
//     operators on u^ (+ W), My, Mx, result in place over its own layer of the u^ plane, bulk-stored from there.  Per dof: 8
//     instead of ~10 shared-memory accesses, 3 instead of 4 hand-offs, no second W round trip.
//   * the XY role is stateless, so its work is cut into TASKS (plane p, layer k) dealt round-robin to NXY warps, NXY free
//     (7 by default): the role that carries 2/3 of the FP64 work is no longer tied to n warps.
//   * no H warps: the Z warps also reduce the x / y halo cells to (V', G') traces of u^ (24 of the 120 halo tasks of a plane
//     per warp; the global loads are issued at the top of the iteration and consumed at its end; Mz is applied to the traces
//     in registers), and the XY warp that finishes the last task of a plane bulk-stores it and refills the stage.
//   * four u stages, two W planes, four halo buffers.  W(q) is free again as soon as the tasks of plane q have READ it (they add
//     Az u half way, after their column sweep), and the Z role hands u^(q) and the halo traces over (ufull) BEFORE it waits for
//     the plane above and finishes W(q) (wfull): the tasks of a plane start without waiting for the next plane to land.
//
// Barriers (mbarrier, counts = producing warps / tasks): ubar[4] plane landed; ufull[4] u^ and halo traces of plane p ready,
// wfull[2] W(p) ready (5 Z warps each); wfree[2] the 5 tasks of plane p have read W; outfull[4] the 5 tasks of plane p are
// done; sfree[4] the bulk store of plane p has read its stage.
#pragma once

template <int n, int NXY_>
struct Roles3 {
  static constexpr int NXY = NXY_, NZW = n, NW = NXY + NZW, NSU = 4, NSW = 2, NSH = 4;
};

// One halo task = the n face points (a, k), k = 0..n-1, of one x-halo (a = j) or y-halo (a = i) face cell of the tile, so that Mz can
// be applied to the traces in registers.  2 TYC n + 2 TXC n = 120 tasks per plane.
struct HaloTask {
  int cell;    // adjacent tile cell (local index at plane z0)
  int off;     // first entry of the line (k = 0) inside the halo cell
  int str;     // stride along the face normal
  int out;     // offset of the pair of face point (a, k = 0) in the halo buffer
  int face;    // 2 dir + side
  int side;
};
template <int n>
__device__ __forceinline__ HaloTask halo_task(const FastGeom &G, int idx, int x0, int y0, int z0) {
  constexpr int NL = n * n;
  const int dir = idx < 2 * TYC * n ? 0 : 1;
  const int e = dir == 0 ? idx : idx - 2 * TYC * n;
  const int per = (dir == 0 ? TYC : TXC) * n;      // tasks per side
  const int side = e / per, rem = e - side * per, cellrow = rem / n, a = rem - cellrow * n;
  const int hcx = dir == 0 ? (side ? TXC - 1 : 0) : cellrow, hcy = dir == 0 ? cellrow : (side ? TYC - 1 : 0);
  HaloTask h;
  h.cell = cell_index(G, x0 + hcx, y0 + hcy, z0);
  h.off = dir == 0 ? a * n : a;
  h.str = dir == 0 ? 1 : n;
  h.out = ((dir == 0 ? 0 : 2 * TYC * NL) + (side * (dir == 0 ? TYC : TXC) + cellrow) * NL + a) * 2;
  h.face = 2 * dir + side; h.side = side;
  return h;
}


// single-line even-odd helpers (the v2 forms take the plain tables as well)
template <int n, typename T>
__device__ __forceinline__ void traces1(const EOParams<T, n> &Q, const T (&u)[n], T &G0, T &G1) {
  constexpr int NE = (n + 1) / 2, NO = n / 2;
  T gs = Q.gE[NE - 1] * ((n & 1) ? u[NE - 1] : (u[NE - 1] + u[n - NE]));
#pragma unroll
  for (int l = 0; l < NE - 1; ++l) gs = fma(Q.gE[l], u[l] + u[n - 1 - l], gs);
  T gd = Q.gO[0] * (u[0] - u[n - 1]);
#pragma unroll
  for (int l = 1; l < NO; ++l) gd = fma(Q.gO[l], u[l] - u[n - 1 - l], gd);
  G0 = gs + gd; G1 = gs - gd;
}
template <int n, typename T>
__device__ __forceinline__ void mass1(const EOParams<T, n> &Q, int d, const T (&q)[n], T (&out)[n]) {
  constexpr int NE = (n + 1) / 2, NO = n / 2;
  T e[NE], o[NO > 0 ? NO : 1];
#pragma unroll
  for (int l = 0; l < NO; ++l) { e[l] = q[l] + q[n - 1 - l]; o[l] = q[l] - q[n - 1 - l]; }
  if (n & 1) e[NE - 1] = q[NE - 1];
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    T s = Q.ME[d][i][0] * e[0];
#pragma unroll
    for (int l = 1; l < NE; ++l) s = fma(Q.ME[d][i][l], e[l], s);
    if ((n & 1) && i == NE - 1) { out[i] = s; continue; }
    T dd = Q.MO[d][i][0] * o[0];
#pragma unroll
    for (int l = 1; l < NO; ++l) dd = fma(Q.MO[d][i][l], o[l], dd);
    out[i] = s + dd; out[n - 1 - i] = s - dd;
  }
}

// top-face (side 1) coefficient vectors of the folded z operator, plain form: the Z warps add this part of Az u once the plane
// above has landed, everything else before
template <typename T, int n>
struct ZTop { T cV[n], cG[n]; };

// even-odd line operator with the side-1 traces absent: out = K' [u; V0', G0', 0, 0] (direction d of EOParams)
template <int n, typename T>
__device__ __forceinline__ void line_eo_side0(const EOParams<T, n> &Q, int d, const T (&u)[n], T V0, T G0, T (&out)[n]) {
  constexpr int NE = (n + 1) / 2, NO = n / 2;
  T e[NE], o[NO > 0 ? NO : 1];
#pragma unroll
  for (int l = 0; l < NO; ++l) { e[l] = u[l] + u[n - 1 - l]; o[l] = u[l] - u[n - 1 - l]; }
  if (n & 1) e[NE - 1] = u[NE - 1];
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    T s = Q.KEv[d][i] * V0;
    s = fma(Q.KEg[d][i], G0, s);
#pragma unroll
    for (int l = 0; l < NE; ++l) s = fma(Q.KE[d][i][l], e[l], s);
    if ((n & 1) && i == NE - 1) { out[i] = s; continue; }
    T dd = Q.KOv[d][i] * V0;
    dd = fma(Q.KOg[d][i], G0, dd);
#pragma unroll
    for (int l = 0; l < NO; ++l) dd = fma(Q.KO[d][i][l], o[l], dd);
    out[i] = s + dd; out[n - 1 - i] = s - dd;
  }
}

// EO tables for v3: direction 2 carries the folded operator Az = (h_z M) [Kt'_z | cV cG]
template <typename T, int n>
void fill_eo_params_v3(EOParams<T, n> &Q, ZTop<T, n> &ZT, const FastParams<double, n> &F) {
  FastParams<double, n> Fz = F;
  for (int i = 0; i < n; ++i) {
    for (int l = 0; l < n; ++l) {
      double a = 0;
      for (int m = 0; m < n; ++m) a += F.Mh[2][i * n + m] * F.Kt[2][m * n + l];
      Fz.Kt[2][i * n + l] = a;
    }
    for (int s = 0; s < 2; ++s) {
      double a = 0, b = 0;
      for (int m = 0; m < n; ++m) { a += F.Mh[2][i * n + m] * F.cV[2][s][m]; b += F.Mh[2][i * n + m] * F.cG[2][s][m]; }
      Fz.cV[2][s][i] = a; Fz.cG[2][s][i] = b;
    }
  }
  fill_eo_params<T, n>(Q, Fz);
  for (int i = 0; i < n; ++i) { ZT.cV[i] = (T)Fz.cV[2][1][i]; ZT.cG[i] = (T)Fz.cG[2][1][i]; }
}

template <int n, typename T, int NXY, int NBX, int NBY, bool ZFIRST = false, int MINB = 1>
__global__ void __launch_bounds__(Roles3<n, NXY>::NW * 32, MINB)
pressure_march3_kernel(const __grid_constant__ EOParams<T, n> Q, const __grid_constant__ ZTop<T, n> ZT, const __grid_constant__ FastGeom G, const int *__restrict__ nbr,
                       const T *__restrict__ src, T *__restrict__ dst, const int *__restrict__ jobs, const __grid_constant__ GhostSrc GS) {
  static_assert(n % 2 == 1, "odd n: the lane stride n^3 must be odd (bank-conflict free without padding)");
  static_assert(NXY >= 1 && NXY <= 2 * n, "a warp's next task lies at most two planes ahead (mbarrier parity waits)");
  using R = Roles3<n, NXY>;
  using P2 = typename Pair<T>::type;
  constexpr int NL = n * n, ND = NL * n, NSU = R::NSU, NSH = R::NSH;
  constexpr int HBUF = (2 * TYC + 2 * TXC) * NL * 2;
  static_assert(NSH == NSU, "halo buffer p % NSH is recycled together with stage p % NSU");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *U = reinterpret_cast<T *>(smem_raw);   // [NSU][TC][ND]  u planes (bulk-copied, memory order) -> u^ = Mz u (Z warps) -> result (XY tasks), all in place
  T *W = U + NSU * TC * ND;                 // [2][TC][ND]    Az u
  T *HXY = W + 2 * TC * ND;                 // [NSH]{[2 side][TYC][NL][2], [2 side][TXC][NL][2]}  (V', G') of u^ on the x / y halo
  unsigned long long *ubar = reinterpret_cast<unsigned long long *>(HXY + NSH * HBUF);   // [NSU]
  unsigned long long *outfull = ubar + NSU;   // [NSU]
  unsigned long long *sfree = outfull + NSU;  // [NSU]
  unsigned long long *ufull = sfree + NSU;    // [NSU]
  unsigned long long *wfull = ufull + NSU;    // [2]
  unsigned long long *wfree = wfull + 2;      // [2]
  int *ucnt = reinterpret_cast<int *>(wfree + 2);   // [NSU] tasks done with the plane held by each stage

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // warps 0 .. NXY-1: XY, then n Z warps; with warp % 4 = sub-partition this spreads both roles over all four
  // (ZFIRST: the Z warps take the low warp ids instead -- the scheduler favours high ids)
  const int role = ZFIRST ? (warp < R::NZW ? 1 : 0) : (warp < R::NXY ? 0 : 1);
  const int ridx = ZFIRST ? (warp < R::NZW ? warp : warp - R::NZW) : (warp < R::NXY ? warp : warp - R::NXY);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSU; ++s) { mbar_init(&ubar[s], 1); mbar_init(&outfull[s], n); mbar_init(&sfree[s], 1); mbar_init(&ufull[s], R::NZW); ucnt[s] = 0; }
    for (int s = 0; s < 2; ++s) { mbar_init(&wfull[s], R::NZW); mbar_init(&wfree[s], n); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int job = jobs ? jobs[blockIdx.x] : blockIdx.x;
  const int tx = job % G.ntx, ty = (job / G.ntx) % G.nty, sz = job / (G.ntx * G.nty);
  const int x0 = G.lox + tx * TXC, y0 = G.loy + ty * TYC, z0 = G.loz + sz * G.NZ;
  const int NZ = G.NZ;
  const size_t no = (size_t)G.n_owned;
  const long long pt0 = plane_term(G, z0);
  // Ghost cells: in the vector tail, or (import still in flight, GS.base != nullptr) in the mailbox slot of this exchange, which is
  // complete once every neighbour's sequence number has arrived.  ghost_wait is called by whole warps.
  bool ghost_ready = GS.base == nullptr;
  auto ghost_wait = [&]() {
    if (ghost_ready) return;
    if (lane < GS.npeers) {
      const unsigned long long *flag = GS.flags + GS.peer_ranks[lane];
      const long long t0 = clock64();
      unsigned long long f;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(flag) : "memory");
        if (clock64() - t0 > 20000000000LL) __trap();   // ~10 s: a lost neighbour must not hang the device
      } while (f < GS.seq);
    }
    __syncwarp();
    ghost_ready = true;
  };
  auto cell_ptr = [&](int idx) -> const T * {
    return (idx < (int)no || GS.base == nullptr) ? src + (size_t)idx * ND : reinterpret_cast<const T *>(GS.base) + (size_t)(idx - (int)no) * ND;
  };

  // called by a whole warp: lane 0 arms the barrier, lanes 0..7 issue one bulk copy of 4 consecutive cells each
  auto issue_plane = [&](int p) {
    unsigned long long *bar = &ubar[p % NSU];
    if (lane == 0) mbar_expect_tx(bar, (unsigned)(TC * ND * sizeof(T)));
    __syncwarp();
    if (lane < TC / 4) {
      const long long sh = plane_term(G, z0 + p) - pt0;
      const int ci = cell_index(G, x0 + slot_cx(4 * lane), y0 + slot_cy(4 * lane), z0);
      bulk_g2s(U + (size_t)(p % NSU) * TC * ND + (size_t)4 * lane * ND, src + (size_t)(ci + sh) * ND, (unsigned)(4 * ND * sizeof(T)), bar);
    }
  };

  if (role == 1) {
    // ------------------------------------------------------------------ Z warps: lane = cell, warp = row j, columns i = 0..n-1
    const int c = lane, j = ridx;
    const int cell0 = cell_index(G, x0 + slot_cx(c), y0 + slot_cy(c), z0);
    const int so = c * ND + j * n;   // shared-memory offset of column (i = 0, j), k = 0
    // per column i: (V', G') of the plane below on the shared face; own top-side derivative of plane q
    T Gt[n], Vt[n], Gown[n];
    {   // ghost plane below the segment: only its top traces are needed
      const int below = nbr[4 * no + cell0];
      if (__any_sync(0xffffffffu, below >= (int)no)) ghost_wait();
      const T *p = cell_ptr(below) + j * n;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T col[n];
#pragma unroll
        for (int k = 0; k < n; ++k) col[k] = __ldg(p + k * NL + i);
        T g0, g1;
        traces1<n, T>(Q, col, g0, g1);
        Gt[i] = g1; Vt[i] = col[n - 1];
      }
    }
    // halo role: 120 tasks per plane = 24 per Z warp (lanes 24..31 redo task 23 and do not store)
    constexpr int HPW = (2 * TYC + 2 * TXC) * n / R::NZW;
    static_assert((2 * TYC + 2 * TXC) * n % R::NZW == 0 && HPW <= 32, "halo tasks per Z warp");
    const HaloTask ht = halo_task<n>(G, j * HPW + (lane < HPW ? lane : HPW - 1), x0, y0, z0);
    // Global loads of this role are issued one iteration (the halo lines) or two (the neighbour index) before their use: the Z
    // role runs a plane ahead of the XY role only if no HBM / L2 latency sits on its critical path.
    T xh[n][n];   // [k][l] halo lines of the plane about to be processed
    auto halo_issue = [&](int hcn) {
      if (__any_sync(0xffffffffu, hcn >= (int)no)) ghost_wait();
      const T *ptr = cell_ptr(hcn) + ht.off;
#pragma unroll
      for (int k = 0; k < n; ++k)
#pragma unroll
        for (int l = 0; l < n; ++l) xh[k][l] = __ldg(ptr + k * NL + l * ht.str);
    };
    auto halo_index = [&](int pl) {   // halo cell of plane pl (clamped to the segment)
      const long long sh = plane_term(G, z0 + (pl < NZ ? pl : NZ - 1)) - pt0;
      return nbr[(size_t)ht.face * no + (ht.cell + sh)];
    };
    halo_issue(halo_index(0));
    int hcn = halo_index(1);
    mbar_wait_sleep(&ubar[0], 0);
#pragma unroll
    for (int i = 0; i < n; ++i) {
      T col[n], g0;
#pragma unroll
      for (int k = 0; k < n; ++k) col[k] = U[so + k * NL + i];
      traces1<n, T>(Q, col, g0, Gown[i]);
    }
    for (int q = 0; q < NZ; ++q) {
      T *Us = U + (size_t)(q % NSU) * TC * ND + so;
      T *Un = U + (size_t)((q + 1) % NSU) * TC * ND + so;
      // a. halo traces of u^ on the n face points (a, k) of my task.  Buffer q % NSH was last read by the tasks of plane q - NSH,
      //    which were done before the refill of stage q % NSU with plane q was issued (NSH == NSU): no wait.
      {
        T *hb = HXY + (size_t)(q % NSH) * HBUF;
        T V[n], Gd[n], Vh[n], Gh[n];
#pragma unroll
        for (int k = 0; k < n; ++k) {
          T g0, g1;
          traces1<n, T>(Q, xh[k], g0, g1);
          V[k] = ht.side ? xh[k][0] : xh[k][n - 1];   // the halo cell's facing side is 1 - side
          Gd[k] = ht.side ? g0 : g1;
        }
        mass1<n, T>(Q, 2, V, Vh);
        mass1<n, T>(Q, 2, Gd, Gh);
        if (lane < HPW) {
#pragma unroll
          for (int k = 0; k < n; ++k) {
            P2 v; v.x = Vh[k]; v.y = Gh[k];
            *reinterpret_cast<P2 *>(hb + ht.out + 2 * n * k) = v;
          }
        }
      }
      // Halo lines of plane q + 1: in flight during step b.  ptxas drains the load scoreboards in front of every wait loop, so
      // their latency is paid in front of the wait of step c -- behind the ufull arrival, off the critical path of the XY role.
      halo_issue(hcn);
      hcn = halo_index(q + 2);
      // b. u^(q) = Mz u over the plane in place (only this thread ever reads these raw columns) and the part of Az u that does
      //    not need the plane above.  The columns of plane q are read back from the stage (not carried in registers: the
      //    register file holds the halo lines in flight and the partial sums instead).
      T acc[n][n];   // [i][k]
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T uc[n], r[n];
#pragma unroll
        for (int k = 0; k < n; ++k) uc[k] = Us[k * NL + i];
        mass1<n, T>(Q, 2, uc, r);
        line_eo_side0<n, T>(Q, 2, uc, Vt[i], Gt[i], acc[i]);
#pragma unroll
        for (int k = 0; k < n; ++k) Us[k * NL + i] = r[k];
        Gt[i] = Gown[i]; Vt[i] = uc[n - 1];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ufull[q % NSU]);   // u^(q) and the halo traces are ready: the tasks of plane q may start
      // c. the plane above
      if (q + 1 < NZ) {
        mbar_wait_sleep(&ubar[(q + 1) % NSU], (unsigned)(((q + 1) / NSU) & 1));
      } else {
        // ghost plane above the segment: stage (q + 1) % NSU is free once the store of plane q - 3 has read it; every thread
        // parks its own columns there and reads them back below
        if (q >= 3) mbar_wait_sleep(&sfree[(q + 1) % NSU], (unsigned)(((q - 3) / NSU) & 1));
        const long long shl = plane_term(G, z0 + NZ - 1) - pt0;
        const int above = nbr[5 * no + (cell0 + shl)];
        if (__any_sync(0xffffffffu, above >= (int)no)) ghost_wait();
        const T *p = cell_ptr(above) + j * n;
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
          for (int k = 0; k < n; ++k) Un[k * NL + i] = __ldg(p + k * NL + i);
      }
      // d. W(q) = Az u: add the top-face terms; the buffer is free once the tasks of plane q - 2 have read it
      if (q >= 2) mbar_wait_sleep(&wfree[q & 1], (unsigned)(((q >> 1) - 1) & 1));
      T *Ws = W + (size_t)(q & 1) * TC * ND + so;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T nw[n], gb, g1n;
#pragma unroll
        for (int k = 0; k < n; ++k) nw[k] = Un[k * NL + i];
        traces1<n, T>(Q, nw, gb, g1n);   // bottom-side derivative of the plane above; its top-side one for later
#pragma unroll
        for (int k = 0; k < n; ++k) Ws[k * NL + i] = fma(ZT.cG[k], gb, fma(ZT.cV[k], nw[0], acc[i][k]));
        Gown[i] = g1n;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&wfull[q & 1]);
    }
  } else {
    // ------------------------------------------------------------------ XY warps: lane = cell, task = (plane p, layer k)
    const int c = lane, cx = slot_cx(c), cy = slot_cy(c);
    if (ridx == 0)
      for (int p = 0; p < NSU && p < NZ; ++p) issue_plane(p);
    const bool ex0 = cx == 0, ex1 = cx == TXC - 1, ey0 = cy == 0, ey1 = cy == TYC - 1;   // tile-edge lanes
    const int lxm = slot_of(ex0 ? cx : cx - 1, cy), lxp = slot_of(ex1 ? cx : cx + 1, cy);
    const int lym = slot_of(cx, ey0 ? cy : cy - 1), lyp = slot_of(cx, ey1 ? cy : cy + 1);
    for (int t = ridx; t < n * NZ; t += R::NXY) {
      const int p = t / n, k = t - p * n, st = p & 1;
      // halo trace pairs of my row / column of the tile edge (lanes inside the tile read pair 0 and keep the shuffled value)
      const int hx0 = ex0 ? ((0 * TYC + cy) * NL + n * k) * 2 : 0, hx1 = ex1 ? ((1 * TYC + cy) * NL + n * k) * 2 : 0;
      const int hy0 = ey0 ? 2 * TYC * NL * 2 + ((0 * TXC + cx) * NL + n * k) * 2 : 0, hy1 = ey1 ? 2 * TYC * NL * 2 + ((1 * TXC + cx) * NL + n * k) * 2 : 0;
      const T *Wp = W + (size_t)st * TC * ND + c * ND + k * NL;
      T *Up = U + (size_t)(p % NSU) * TC * ND + c * ND + k * NL;
      const T *hb = HXY + (size_t)(p % NSH) * HBUF;
      mbar_wait_sleep(&ufull[p % NSU], (unsigned)((p / NSU) & 1));
      T tl[n][n];   // [row j][column i] of the layer: W + Lx u^, then + Ly u^, then My, then Mx
      auto kbatch = [&](auto isrow_c, auto a0_c, auto nb_c) {
        constexpr bool isrow = decltype(isrow_c)::value;
        constexpr int a0 = decltype(a0_c)::value, NB = decltype(nb_c)::value;
        T u[NB][n], out[NB][n], g0[NB], g1[NB], nv[NB][4];
        P2 h0[NB], h1[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
          for (int l = 0; l < n; ++l) u[b][l] = isrow ? Up[(a0 + b) * n + l] : Up[l * n + a0 + b];
          if constexpr (!isrow) {
#pragma unroll
            for (int l = 0; l < n; ++l) out[b][l] = tl[l][a0 + b];
          }
          h0[b] = *reinterpret_cast<const P2 *>(hb + (isrow ? hx0 : hy0) + 2 * (a0 + b));
          h1[b] = *reinterpret_cast<const P2 *>(hb + (isrow ? hx1 : hy1) + 2 * (a0 + b));
        }
        traces_eo_nb<n, T, NB>(Q, u, g0, g1);
        const int lm = isrow ? lxm : lym, lp = isrow ? lxp : lyp;
        const bool e0 = isrow ? ex0 : ey0, e1 = isrow ? ex1 : ey1;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const T v0 = __shfl_sync(0xffffffffu, u[b][n - 1], lm), w0 = __shfl_sync(0xffffffffu, g1[b], lm);
          const T v1 = __shfl_sync(0xffffffffu, u[b][0], lp), w1 = __shfl_sync(0xffffffffu, g0[b], lp);
          nv[b][0] = e0 ? h0[b].x : v0; nv[b][1] = e0 ? h0[b].y : w0;
          nv[b][2] = e1 ? h1[b].x : v1; nv[b][3] = e1 ? h1[b].y : w1;
        }
        lines_eo_nb<n, T, NB, !isrow>(Q, isrow ? 0 : 1, u, nv, out);
        if constexpr (isrow) {
#pragma unroll
          for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int l = 0; l < n; ++l) tl[a0 + b][l] = out[b][l];
        } else {
          // Az u joins here: the only read of W, half way through the task (the Z warps finish W(p) after u^(p))
          if constexpr (a0 == 0) mbar_wait_sleep(&wfull[st], (unsigned)((p >> 1) & 1));
#pragma unroll
          for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int l = 0; l < n; ++l) out[b][l] += Wp[l * n + a0 + b];
          T res[NB][n];
          mass_eo_nb<n, T, NB>(Q, 1, out, res);   // My
#pragma unroll
          for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int l = 0; l < n; ++l) tl[l][a0 + b] = res[b][l];
        }
      };
      auto xmass = [&](auto a0_c, auto nb_c) {
        constexpr int a0 = decltype(a0_c)::value, NB = decltype(nb_c)::value;
        T q[NB][n], res[NB][n];
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int l = 0; l < n; ++l) q[b][l] = tl[a0 + b][l];
        mass_eo_nb<n, T, NB>(Q, 0, q, res);       // Mx
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int l = 0; l < n; ++l) Up[(a0 + b) * n + l] = res[b][l];   // over my own layer of u^: no other task reads it
      };
      // batches of NBX rows / NBY columns in lockstep (coefficient-stationary), remainder batch last
      auto sweep = [&](auto isrow_c, auto nb_c) {
        constexpr int NB = decltype(nb_c)::value;
        [&]<int... I>(std::integer_sequence<int, I...>) { (kbatch(isrow_c, std::integral_constant<int, I * NB>{}, nb_c), ...); }(std::make_integer_sequence<int, n / NB>{});
        if constexpr (n % NB != 0) kbatch(isrow_c, std::integral_constant<int, (n / NB) * NB>{}, std::integral_constant<int, n % NB>{});
      };
      sweep(std::true_type{}, std::integral_constant<int, NBX>{});
      sweep(std::false_type{}, std::integral_constant<int, NBY>{});
      // W(p) is read: the Z warps may overwrite it with plane p + 2
      __syncwarp();
      if (lane == 0) mbar_arrive(&wfree[st]);
      {
        constexpr int NB = NBX;
        [&]<int... I>(std::integer_sequence<int, I...>) { (xmass(std::integral_constant<int, I * NB>{}, std::integral_constant<int, NB>{}), ...); }(std::make_integer_sequence<int, n / NB>{});
        if constexpr (n % NB != 0) xmass(std::integral_constant<int, (n / NB) * NB>{}, std::integral_constant<int, n % NB>{});
      }
      // the task is done: its layer of the stage is final (the bulk store reads it through the async proxy).  The warp that
      // finishes the fifth task of a plane stores the plane and then refills the stage with plane p + NSU.
      fence_proxy_async();
      __syncwarp();
      int last = 0;
      if (lane == 0) {
        mbar_arrive(&outfull[p % NSU]);
        const int old = atomicAdd(&ucnt[p % NSU], 1);
        if (old == n - 1) { ucnt[p % NSU] = 0; last = 1; }
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        mbar_wait_sleep(&outfull[p % NSU], (unsigned)((p / NSU) & 1));   // orders the other warps' writes before the bulk read
        if (lane < TC / 4) {
          const long long sh = plane_term(G, z0 + p) - pt0;
          const int ci = cell_index(G, x0 + slot_cx(4 * lane), y0 + slot_cy(4 * lane), z0);
          bulk_s2g(dst + (size_t)(ci + sh) * ND, U + (size_t)(p % NSU) * TC * ND + (size_t)4 * lane * ND, (unsigned)(4 * ND * sizeof(T)));
          bulk_commit();
          bulk_wait_read0();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sfree[p % NSU]);
        if (p + NSU < NZ) issue_plane(p + NSU);
      }
    }
    if (lane < TC / 4) bulk_wait0();
  }
}
