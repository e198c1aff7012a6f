// Second-generation marching kernel of the pressure operator (odd n; instantiated for Q4 = the headline workload).
// Included by pressure_fast.cu inside its anonymous namespace (uses FastParams, FastGeom, the bulk-copy helpers).
//
// Same algebra and the same three roles as pressure_march_kernel (see the header of pressure_fast.cu); what changed is
// how the roles meet, after the round-1 ncu capture showed a latency / hand-off bound pipeline (barrier stall 1.41 per
// issue, 30 % of the shared-memory wavefronts bank conflicts, one CTA of 12 warps per SM):
//
//   * every hand-off is an mbarrier that counts ONLY the producing warps (full: Z + H warps -> XY; qready: XY -> Z, H;
//     outfull / outfree: Z -> store issuer -> Z).  No role waits for warps of its own role any more; the XY warps have no
//     barrier among themselves at all: the x / y neighbours of a cell's layer live in the SAME warp (lane = cell of the
//     8 x 4 tile plane, warp = layer), so the outward normal derivative traces travel by warp shuffle and only the lanes
//     on the tile edge read the halo traces from shared memory.
//   * Z warps: lane = cell, warp = row j, 5 columns (i = 0..4) per thread.  With n^3 = 125 (odd) the lane stride 125 hits
//     16 distinct 8-byte banks per half warp: every shared-memory access of every role is conflict free, no padded copies.
//   * the result leaves through the TMA engine as well: Z warps write Mz q into a staging plane in memory order, one H
//     warp issues 8 bulk stores of 4 consecutive cells (cp.async.bulk.global.shared::cta) per plane.  No STG in the march.
//   * optional even-odd form of all 1-D contractions (EO): the folded line operator [Kt' | cV cG] (n x (n + 4)) and h M are
//     centro-symmetric, so a line costs 2 (n/2)-blocks instead of one n-block: 48 -> ~42 FP64 instructions per dof.
#pragma once

template <int n>
struct Roles2 {
  static constexpr int NXY = n, NZW = n, NH = 2, NW = NXY + NZW + NH, NSTAGE = 3;
};

// Wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or ~1 ms passed) instead of
// polling -- the polling loops of the waiting roles were 32 % of all executed instructions of the first v2 capture and
// competed with the XY warps for issue slots.
// spin (mode != 0) or sleep (mode == 0): measured choice per role, see launch_march2
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long *bar, unsigned parity);
__device__ __forceinline__ void mbar_wait_mode(unsigned long long *bar, unsigned parity, int spin) {
  if (spin) mbar_wait(bar, parity);
  else mbar_wait_sleep(bar, parity);
}
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long *bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  unsigned ok = 0;
  long long t0 = 0;
  for (int spin = 0; !ok; ++spin) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(addr), "r"(parity), "r"(1000000u) : "memory");
    // never hang the device on a lost arrival -- but by the clock, not by the poll count: with the ghost import in flight a CTA
    // may legitimately wait for a neighbour rank that is still busy (e.g. uploading its input over PCIe) for many milliseconds
    if ((spin & 1023) == 1023) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 40000000000LL) __trap();   // ~20 s
    }
  }
}
// (V', G') pair from shared memory for the lanes with pred != 0, the others keep their values: one predicated LDS.128, no
// branch (an `if` around the load made the compiler fork the warp and run the line operators once per lane group)
__device__ __forceinline__ void lds_pair_if(double &v, double &w, const double *p, int pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q ld.shared.v2.f64 {%0, %1}, [%2]; }" : "+d"(v), "+d"(w) : "r"(smem_u32(p)), "r"(pred) : "memory");
}
__device__ __forceinline__ void lds_pair_if(float &v, float &w, const float *p, int pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q ld.shared.v2.f32 {%0, %1}, [%2]; }" : "+f"(v), "+f"(w) : "r"(smem_u32(p)), "r"(pred) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bulk async copy shared -> global (TMA engine), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T> struct Pair;
template <> struct Pair<double> { using type = double2; };
template <> struct Pair<float> { using type = float2; };

// Even-odd tables of one direction (n odd or even): NE = ceil(n/2) even outputs, NO = floor(n/2) odd outputs.
//   S_i = sum_l KE[i][l] e_l + KEv[i] (V0' + V1') + KEg[i] (G0' + G1'),   e_l = u_l + u_{n-1-l} (middle: u_l)
//   D_i = sum_l KO[i][l] o_l + KOv[i] (V0' - V1') + KOg[i] (G0' - G1'),   o_l = u_l - u_{n-1-l}
//   out_i = S_i + D_i,  out_{n-1-i} = S_i - D_i
template <typename T, int n>
struct EOParams {
  static constexpr int NE = (n + 1) / 2, NO = n / 2;
  T KE[3][NE][NE], KO[3][NO > 0 ? NO : 1][NO > 0 ? NO : 1];
  T KEv[3][NE], KEg[3][NE], KOv[3][NO > 0 ? NO : 1], KOg[3][NO > 0 ? NO : 1];
  T ME[3][NE][NE], MO[3][NO > 0 ? NO : 1][NO > 0 ? NO : 1];
  T gE[NE], gO[NO > 0 ? NO : 1];   // G_0 = gs + gd, G_1 = gs - gd with gs = sum gE e, gd = sum gO o
};

// ---- 1-D line operators (registers only) ----------------------------------------------------------------------------
// plain form: out[i] (+)= sum_l K[i][l] u[l] + cV0[i] V0 + cG0[i] G0 + cV1[i] V1 + cG1[i] G1
template <int n, typename T, bool ACC>
__device__ __forceinline__ void line_plain(const T *K, const T *cV0, const T *cG0, const T *cV1, const T *cG1, const T (&u)[n], T V0, T G0, T V1,
                                           T G1, T (&out)[n]) {
#pragma unroll
  for (int i = 0; i < n; ++i) {
    T a = ACC ? fma(cV0[i], V0, out[i]) : cV0[i] * V0;
    a = fma(cG0[i], G0, a);
    a = fma(cV1[i], V1, a);
    a = fma(cG1[i], G1, a);
#pragma unroll
    for (int l = 0; l < n; ++l) a = fma(K[i * n + l], u[l], a);
    out[i] = a;
  }
}
// even-odd form of the same line operator (direction d of EOParams)
template <int n, typename T, bool ACC>
__device__ __forceinline__ void line_eo(const EOParams<T, n> &Q, int d, const T (&u)[n], T V0, T G0, T V1, T G1, T (&out)[n]) {
  constexpr int NE = (n + 1) / 2, NO = n / 2;
  T e[NE], o[NO > 0 ? NO : 1];
#pragma unroll
  for (int l = 0; l < NO; ++l) { e[l] = u[l] + u[n - 1 - l]; o[l] = u[l] - u[n - 1 - l]; }
  if (n & 1) e[NE - 1] = u[NE - 1];
  const T ve = V0 + V1, vo = V0 - V1, ge = G0 + G1, go = G0 - G1;
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    const bool mid = (n & 1) && i == NE - 1;
    T s = (ACC && mid) ? fma(Q.KEv[d][i], ve, out[i]) : Q.KEv[d][i] * ve;
    s = fma(Q.KEg[d][i], ge, s);
#pragma unroll
    for (int l = 0; l < NE; ++l) s = fma(Q.KE[d][i][l], e[l], s);
    if (mid) { out[i] = s; continue; }
    T dd = Q.KOv[d][i] * vo;
    dd = fma(Q.KOg[d][i], go, dd);
#pragma unroll
    for (int l = 0; l < NO; ++l) dd = fma(Q.KO[d][i][l], o[l], dd);
    if (ACC) { out[i] = (out[i] + s) + dd; out[n - 1 - i] = (out[n - 1 - i] + s) - dd; }
    else { out[i] = s + dd; out[n - 1 - i] = s - dd; }
  }
}
// mass line: out = (h M) q
template <int n, typename T, bool EO>
__device__ __forceinline__ void line_mass(const FastParams<T, n> &P, const EOParams<T, n> &Q, int d, const T (&q)[n], T (&out)[n]) {
  if constexpr (EO) {
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
  } else {
#pragma unroll
    for (int i = 0; i < n; ++i) {
      T a = P.Mh[d][i * n] * q[0];
#pragma unroll
      for (int l = 1; l < n; ++l) a = fma(P.Mh[d][i * n + l], q[l], a);
      out[i] = a;
    }
  }
}
// own outward normal derivative traces of a line: G0 (side 0), G1 (side 1)
template <int n, typename T, bool EO>
__device__ __forceinline__ void line_traces(const FastParams<T, n> &P, const EOParams<T, n> &Q, const T (&u)[n], T &G0, T &G1) {
  if constexpr (EO) {
    constexpr int NE = (n + 1) / 2, NO = n / 2;
    T gs = Q.gE[NE - 1] * ((n & 1) ? u[NE - 1] : (u[NE - 1] + u[n - NE]));
#pragma unroll
    for (int l = 0; l < NE - 1; ++l) gs = fma(Q.gE[l], u[l] + u[n - 1 - l], gs);
    T gd = Q.gO[0] * (u[0] - u[n - 1]);
#pragma unroll
    for (int l = 1; l < NO; ++l) gd = fma(Q.gO[l], u[l] - u[n - 1 - l], gd);
    G0 = gs + gd; G1 = gs - gd;
  } else {
    T a = P.g[0][0] * u[0], b = P.g[1][0] * u[0];
#pragma unroll
    for (int l = 1; l < n; ++l) { a = fma(P.g[0][l], u[l], a); b = fma(P.g[1][l], u[l], b); }
    G0 = a; G1 = b;
  }
}

// ---- NB lines in lockstep (coefficient-stationary): every table entry is fetched once and feeds NB independent FMAs, so a
// thread always has NB FP64 chains in flight (DFMA latency is ~9 cycles at one issue per ~2.3 cycles, tools/fp64_issue.cu)
// and the constant-bank loads (Blackwell FP64 instructions take no c[] operand: LDC + register) are amortised over NB lines.
template <int n, typename T, int NB>
__device__ __forceinline__ void traces_eo_nb(const EOParams<T, n> &Q, const T (&u)[NB][n], T (&G0)[NB], T (&G1)[NB]) {
  constexpr int NE = (n + 1) / 2, NO = n / 2;
  T gs[NB], gd[NB];
  {
    const T c = Q.gE[NE - 1];
#pragma unroll
    for (int b = 0; b < NB; ++b) gs[b] = c * ((n & 1) ? u[b][NE - 1] : (u[b][NE - 1] + u[b][n - NE]));
  }
#pragma unroll
  for (int l = 0; l < NE - 1; ++l) {
    const T c = Q.gE[l];
#pragma unroll
    for (int b = 0; b < NB; ++b) gs[b] = fma(c, u[b][l] + u[b][n - 1 - l], gs[b]);
  }
#pragma unroll
  for (int l = 0; l < NO; ++l) {
    const T c = Q.gO[l];
#pragma unroll
    for (int b = 0; b < NB; ++b) gd[b] = l == 0 ? c * (u[b][l] - u[b][n - 1 - l]) : fma(c, u[b][l] - u[b][n - 1 - l], gd[b]);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) { G0[b] = gs[b] + gd[b]; G1[b] = gs[b] - gd[b]; }
}
// out[b] += K' [u[b]; V0', G0', V1', G1'] for NB lines of direction d (nv[b] = V0', G0', V1', G1')
template <int n, typename T, int NB, bool ACC = true>
__device__ __forceinline__ void lines_eo_nb(const EOParams<T, n> &Q, const int d, const T (&u)[NB][n], const T (&nv)[NB][4], T (&out)[NB][n]) {
  constexpr int NE = (n + 1) / 2, NO = n / 2;
  T e[NB][NE], o[NB][NO > 0 ? NO : 1], ve[NB], vo[NB], ge[NB], go[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
#pragma unroll
    for (int l = 0; l < NO; ++l) { e[b][l] = u[b][l] + u[b][n - 1 - l]; o[b][l] = u[b][l] - u[b][n - 1 - l]; }
    if (n & 1) e[b][NE - 1] = u[b][NE - 1];
    ve[b] = nv[b][0] + nv[b][2]; vo[b] = nv[b][0] - nv[b][2]; ge[b] = nv[b][1] + nv[b][3]; go[b] = nv[b][1] - nv[b][3];
  }
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    const bool mid = (n & 1) && i == NE - 1;
    T s[NB], dd[NB];
    {
      const T c = Q.KEv[d][i];
#pragma unroll
      for (int b = 0; b < NB; ++b) s[b] = (mid && ACC) ? fma(c, ve[b], out[b][i]) : c * ve[b];
    }
    {
      const T c = Q.KEg[d][i];
#pragma unroll
      for (int b = 0; b < NB; ++b) s[b] = fma(c, ge[b], s[b]);
    }
#pragma unroll
    for (int l = 0; l < NE; ++l) {
      const T c = Q.KE[d][i][l];
#pragma unroll
      for (int b = 0; b < NB; ++b) s[b] = fma(c, e[b][l], s[b]);
    }
    if (mid) {
#pragma unroll
      for (int b = 0; b < NB; ++b) out[b][i] = s[b];
      continue;
    }
    {
      const T c = Q.KOv[d][i];
#pragma unroll
      for (int b = 0; b < NB; ++b) dd[b] = c * vo[b];
    }
    {
      const T c = Q.KOg[d][i];
#pragma unroll
      for (int b = 0; b < NB; ++b) dd[b] = fma(c, go[b], dd[b]);
    }
#pragma unroll
    for (int l = 0; l < NO; ++l) {
      const T c = Q.KO[d][i][l];
#pragma unroll
      for (int b = 0; b < NB; ++b) dd[b] = fma(c, o[b][l], dd[b]);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      if (ACC) { out[b][i] = (out[b][i] + s[b]) + dd[b]; out[b][n - 1 - i] = (out[b][n - 1 - i] + s[b]) - dd[b]; }
      else { out[b][i] = s[b] + dd[b]; out[b][n - 1 - i] = s[b] - dd[b]; }
    }
  }
}
// out[b] = (h M) q[b] for NB lines of direction d
template <int n, typename T, int NB>
__device__ __forceinline__ void mass_eo_nb(const EOParams<T, n> &Q, const int d, const T (&q)[NB][n], T (&out)[NB][n]) {
  constexpr int NE = (n + 1) / 2, NO = n / 2;
  T e[NB][NE], o[NB][NO > 0 ? NO : 1];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
#pragma unroll
    for (int l = 0; l < NO; ++l) { e[b][l] = q[b][l] + q[b][n - 1 - l]; o[b][l] = q[b][l] - q[b][n - 1 - l]; }
    if (n & 1) e[b][NE - 1] = q[b][NE - 1];
  }
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    const bool mid = (n & 1) && i == NE - 1;
    T s[NB], dd[NB];
#pragma unroll
    for (int l = 0; l < NE; ++l) {
      const T c = Q.ME[d][i][l];
#pragma unroll
      for (int b = 0; b < NB; ++b) s[b] = l == 0 ? c * e[b][l] : fma(c, e[b][l], s[b]);
    }
    if (mid) {
#pragma unroll
      for (int b = 0; b < NB; ++b) out[b][i] = s[b];
      continue;
    }
#pragma unroll
    for (int l = 0; l < NO; ++l) {
      const T c = Q.MO[d][i][l];
#pragma unroll
      for (int b = 0; b < NB; ++b) dd[b] = l == 0 ? c * o[b][l] : fma(c, o[b][l], dd[b]);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) { out[b][i] = s[b] + dd[b]; out[b][n - 1 - i] = s[b] - dd[b]; }
  }
}

template <int n, typename T, bool EO, int XYV>
__global__ void __launch_bounds__(Roles2<n>::NW * 32, 1)
pressure_march2_kernel(const __grid_constant__ FastParams<T, n> P, const __grid_constant__ EOParams<T, n> Q, const __grid_constant__ FastGeom G,
                       const int *__restrict__ nbr, const T *__restrict__ src, T *__restrict__ dst, const int *__restrict__ jobs) {
  static_assert(n % 2 == 1, "odd n: the lane stride n^3 must be odd (bank-conflict free without padding)");
  using R = Roles2<n>;
  using P2 = typename Pair<T>::type;
  constexpr int NL = n * n, ND = NL * n, NS = R::NSTAGE;
  constexpr int HBUF = (2 * TYC + 2 * TXC) * NL * 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *U = reinterpret_cast<T *>(smem_raw);   // [NS][TC][ND]  u planes, bulk-copied, memory order
  T *W = U + NS * TC * ND;                  // [2][TC][ND]   z-partial sums, then the xy result (in place)
  T *OUT = W + 2 * TC * ND;                 // [TC][ND]      finished plane, memory order, bulk-stored
  T *HXY = OUT + TC * ND;                   // [2]{[2 side][TYC][NL][2], [2 side][TXC][NL][2]}  (V', G') of the x / y halo
  unsigned long long *ubar = reinterpret_cast<unsigned long long *>(HXY + 2 * HBUF);   // [NS] plane landed
  unsigned long long *full = ubar + NS;     // [2] W + halo traces of plane p ready      (Z + H warps arrive)
  unsigned long long *qready = full + 2;    // [2] xy result of plane p in W             (XY warps arrive)
  unsigned long long *outfull = qready + 2; // OUT holds plane p                          (Z warps arrive)
  unsigned long long *outfree = outfull + 1;// the bulk stores of plane p have read OUT  (issuer arrives)
  int *ucnt = reinterpret_cast<int *>(outfree + 1);   // [NS] XY warps done reading the plane held by each stage

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // roles interleaved over the 4 sub-partitions (warp % 4): XY warps carry twice the FP64 work of Z warps
  int role, ridx;   // 0 = XY, 1 = Z, 2 = H
  {
    static_assert(n == 5, "role table written for n = 5");
    const int role_tab[12] = {0, 0, 0, 1, 0, 0, 1, 1, 2, 2, 1, 1};
    const int idx_tab[12] = {0, 2, 4, 2, 1, 3, 0, 3, 0, 1, 1, 4};
    role = role_tab[warp]; ridx = idx_tab[warp];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&ubar[s], 1); ucnt[s] = 0; }
    for (int s = 0; s < 2; ++s) { mbar_init(&full[s], R::NZW + R::NH); mbar_init(&qready[s], R::NXY); }
    mbar_init(outfull, R::NZW);
    mbar_init(outfree, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int job = jobs ? jobs[blockIdx.x] : blockIdx.x;
  const int tx = job % G.ntx, ty = (job / G.ntx) % G.nty, sz = job / (G.ntx * G.nty);
  const int x0 = G.lox + tx * TXC, y0 = G.loy + ty * TYC, z0 = G.loz + sz * G.NZ;
  const int NZ = G.NZ;
  const size_t no = (size_t)G.n_owned;
  const long long pt0 = plane_term(G, z0);

  // called by a whole warp: lane 0 arms the barrier, lanes 0..7 issue one bulk copy of 4 consecutive cells each
  auto issue_plane = [&](int p) {
    unsigned long long *bar = &ubar[p % NS];
    if (lane == 0) mbar_expect_tx(bar, (unsigned)(TC * ND * sizeof(T)));
    __syncwarp();
    if (lane < TC / 4) {
      const long long sh = plane_term(G, z0 + p) - pt0;
      const int ci = cell_index(G, x0 + slot_cx(4 * lane), y0 + slot_cy(4 * lane), z0);
      bulk_g2s(U + (size_t)(p % NS) * TC * ND + (size_t)4 * lane * ND, src + (size_t)(ci + sh) * ND, (unsigned)(4 * ND * sizeof(T)), bar);
    }
  };

  if (role == 1) {
    // ------------------------------------------------------------------ Z warps: lane = cell, warp = row j, columns i = 0..n-1
    const int c = lane, j = ridx;
    const int cell0 = cell_index(G, x0 + slot_cx(c), y0 + slot_cy(c), z0);   // local cell index at plane z0
    const int so = c * ND + j * n;                                           // shared-memory offset of column (i = 0, j), k = 0
    // [i][k]: plane q of my columns; (V', G') of the plane below on the shared face; own top-side derivative of plane q
    T uc[n][n], Gt[n], Vt[n], Gown[n];
    {   // ghost plane below the segment: only its top traces are needed
      const T *p = src + (size_t)nbr[4 * no + cell0] * ND + j * n;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T col[n];
#pragma unroll
        for (int k = 0; k < n; ++k) col[k] = __ldg(p + k * NL + i);
        T g0, g1;
        line_traces<n, T, false>(P, Q, col, g0, g1);
        Gt[i] = g1; Vt[i] = col[n - 1];
      }
    }
    auto mz_store = [&](int p) {   // OUT(plane p) = Mz Mx q
      mbar_wait_mode(&qready[p & 1], (unsigned)((p >> 1) & 1), G.wait_mode & 2);
      mbar_wait_mode(outfree, (unsigned)((p - 1) & 1), G.wait_mode & 2);   // the stores of plane p - 1 have read OUT (passes at once for p = 0)
      const T *Ws = W + (size_t)(p & 1) * TC * ND + so;
      T q[n][n];   // [k][i]: my row of every layer
#pragma unroll
      for (int k = 0; k < n; ++k) {
        T row[n];
#pragma unroll
        for (int i = 0; i < n; ++i) row[i] = Ws[k * NL + i];
        line_mass<n, T, EO>(P, Q, 0, row, q[k]);      // Mx
      }
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T col[n], r[n];
#pragma unroll
        for (int k = 0; k < n; ++k) col[k] = q[k][i];
        line_mass<n, T, EO>(P, Q, 2, col, r);         // Mz
#pragma unroll
        for (int k = 0; k < n; ++k) OUT[so + k * NL + i] = r[k];
      }
      fence_proxy_async();   // generic-proxy writes before the async-proxy (bulk store) reads
      __syncwarp();
      if (lane == 0) mbar_arrive(outfull);
    };
    mbar_wait_mode(&ubar[0], 0, G.wait_mode & 2);
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int k = 0; k < n; ++k) uc[i][k] = U[so + k * NL + i];
      T g0;
      line_traces<n, T, EO>(P, Q, uc[i], g0, Gown[i]);
    }
    for (int q = 0; q < NZ; ++q) {
      T *Un = U + (size_t)((q + 1) % NS) * TC * ND + so;
      if (q + 1 < NZ) {
        mbar_wait_mode(&ubar[(q + 1) % NS], (unsigned)(((q + 1) / NS) & 1), G.wait_mode & 2);
      } else {
        // ghost plane above the segment: stage (q + 1) % NS is free (plane q - 2 went through the XY warps before the last
        // mz_store); every thread parks its own columns there and reads them back through the common path
        const long long shl = plane_term(G, z0 + NZ - 1) - pt0;
        const T *p = src + (size_t)nbr[5 * no + (cell0 + shl)] * ND + j * n;
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
          for (int k = 0; k < n; ++k) Un[k * NL + i] = __ldg(p + k * NL + i);
      }
      T *Ws = W + (size_t)(q & 1) * TC * ND + so;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T nw[n], gb, g1n, acc[n];
#pragma unroll
        for (int k = 0; k < n; ++k) nw[k] = Un[k * NL + i];
        line_traces<n, T, EO>(P, Q, nw, gb, g1n);   // bottom-side derivative of the plane above; its top-side one for later
        if constexpr (EO) line_eo<n, T, false>(Q, 2, uc[i], Vt[i], Gt[i], nw[0], gb, acc);
        else line_plain<n, T, false>(P.Kt[2], P.cV[2][0], P.cG[2][0], P.cV[2][1], P.cG[2][1], uc[i], Vt[i], Gt[i], nw[0], gb, acc);
#pragma unroll
        for (int k = 0; k < n; ++k) Ws[k * NL + i] = acc[k];
        Gt[i] = Gown[i]; Vt[i] = uc[i][n - 1]; Gown[i] = g1n;
#pragma unroll
        for (int k = 0; k < n; ++k) uc[i][k] = nw[k];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[q & 1]);
      if (q >= 1) mz_store(q - 1);
    }
    mz_store(NZ - 1);
  } else if (role == 0) {
    // ------------------------------------------------------------------ XY warps: lane = cell, warp = layer k
    const int k = ridx, c = lane, cx = slot_cx(c), cy = slot_cy(c);
    if (k == 0)
      for (int p = 0; p < NS && p < NZ; ++p) issue_plane(p);
    const bool ex0 = cx == 0, ex1 = cx == TXC - 1, ey0 = cy == 0, ey1 = cy == TYC - 1;   // tile-edge lanes
    const int lxm = slot_of(ex0 ? cx : cx - 1, cy), lxp = slot_of(ex1 ? cx : cx + 1, cy);
    const int lym = slot_of(cx, ey0 ? cy : cy - 1), lyp = slot_of(cx, ey1 ? cy : cy + 1);
    // halo trace pairs (V', G') of my row / column of the tile edge, relative to the plane's halo buffer
    // (lanes inside the tile read pair 0 of the buffer and keep the shuffled value: plain loads, no predication, no branch)
    const int hx0 = ex0 ? ((0 * TYC + cy) * NL + n * k) * 2 : 0, hx1 = ex1 ? ((1 * TYC + cy) * NL + n * k) * 2 : 0;
    const int hy0 = ey0 ? 2 * TYC * NL * 2 + ((0 * TXC + cx) * NL + n * k) * 2 : 0, hy1 = ey1 ? 2 * TYC * NL * 2 + ((1 * TXC + cx) * NL + n * k) * 2 : 0;
    for (int p = 0; p < NZ; ++p) {
      const int st = p & 1;
      T *Wp = W + (size_t)st * TC * ND + c * ND + k * NL;
      const T *Up = U + (size_t)(p % NS) * TC * ND + c * ND + k * NL;
      const T *hb = HXY + (size_t)st * HBUF;
      mbar_wait_mode(&ubar[p % NS], (unsigned)((p / NS) & 1), G.wait_mode & 1);
      if constexpr (XYV == 2) {
        // Lockstep form (even-odd only): the n rows, then the n columns of the layer go through the batched line operators
        // in groups of 3 + 2 lines; only the x results (xr) stay in registers between the two passes.
        static_assert(XYV != 2 || EO, "the lockstep form is written for the even-odd tables");
        mbar_wait_mode(&full[st], (unsigned)((p >> 1) & 1), G.wait_mode & 1);
        T xr[n][n];
        auto batch = [&](auto isrow_c, auto a0_c, auto nb_c) {
          constexpr bool isrow = decltype(isrow_c)::value;
          constexpr int a0 = decltype(a0_c)::value, NB = decltype(nb_c)::value;
          T u[NB][n], out[NB][n], g0[NB], g1[NB], nv[NB][4];
          P2 h0[NB], h1[NB];
#pragma unroll
          for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int l = 0; l < n; ++l) u[b][l] = isrow ? Up[(a0 + b) * n + l] : Up[l * n + a0 + b];
#pragma unroll
            for (int l = 0; l < n; ++l) out[b][l] = isrow ? Wp[(a0 + b) * n + l] : xr[l][a0 + b];
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
          lines_eo_nb<n, T, NB>(Q, isrow ? 0 : 1, u, nv, out);
          if constexpr (isrow) {
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
              for (int l = 0; l < n; ++l) xr[a0 + b][l] = out[b][l];
          } else {
            T res[NB][n];
            mass_eo_nb<n, T, NB>(Q, 1, out, res);   // mass in y (the Z warps apply Mx and Mz)
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
              for (int l = 0; l < n; ++l) Wp[l * n + a0 + b] = res[b][l];
          }
        };
        constexpr int NB1 = (n + 1) / 2, NB2 = n - NB1;
        batch(std::true_type{}, std::integral_constant<int, 0>{}, std::integral_constant<int, NB1>{});
        batch(std::true_type{}, std::integral_constant<int, NB1>{}, std::integral_constant<int, NB2>{});
        batch(std::false_type{}, std::integral_constant<int, 0>{}, std::integral_constant<int, NB1>{});
        batch(std::false_type{}, std::integral_constant<int, NB1>{}, std::integral_constant<int, NB2>{});
        // all reads of the stage are done (the W stores above depend on them): the last XY warp refills it
        {
          __syncwarp();
          int last = 0;
          if (lane == 0) {
            const int old = atomicAdd(&ucnt[p % NS], 1);
            if (old == R::NXY - 1) { ucnt[p % NS] = 0; last = 1; }
          }
          last = __shfl_sync(0xffffffffu, last, 0);
          if (last && p + NS < NZ) issue_plane(p + NS);
        }
      } else if constexpr (XYV == 1) {
        // Streaming form: only the x results stay in registers (xr); the 2 n lines of the layer (n rows, then n columns) are
        // read from the stage when needed and go through an explicit 3-stage software pipeline, so that the loads of line
        // t + 2 and the trace exchange of line t + 1 are in flight while line t is in the FP64 pipe:
        //   A(t + 2): u line, W row (rows only), halo pairs        B(t + 1): own traces, shuffles, tile-edge select
        //   C(t):     line operator (+ mass in y and the store for columns)
        mbar_wait_mode(&full[st], (unsigned)((p >> 1) & 1), G.wait_mode & 1);
        T xr[n][n];
        T ln[3][n], wr[3][n];       // u line / W row of the lines in flight
        P2 hp[3][2];                // halo pairs (V', G') of the two sides
        T nv[2][4];                 // neighbour traces v0, w0, v1, w1 of the lines in stages B / C
        auto stageA = [&](int t, int s) {
          const bool isrow = t < n;
          const int a = isrow ? t : t - n;
#pragma unroll
          for (int l = 0; l < n; ++l) ln[s][l] = isrow ? Up[a * n + l] : Up[l * n + a];
          if (isrow) {
#pragma unroll
            for (int l = 0; l < n; ++l) wr[s][l] = Wp[a * n + l];
          }
          hp[s][0] = *reinterpret_cast<const P2 *>(hb + (isrow ? hx0 : hy0) + 2 * a);
          hp[s][1] = *reinterpret_cast<const P2 *>(hb + (isrow ? hx1 : hy1) + 2 * a);
        };
        auto stageB = [&](int t, int s, int s2) {
          const bool isrow = t < n;
          T g0, g1;
          line_traces<n, T, EO>(P, Q, ln[s], g0, g1);
          const int lm = isrow ? lxm : lym, lp = isrow ? lxp : lyp;
          const bool e0 = isrow ? ex0 : ey0, e1 = isrow ? ex1 : ey1;
          const T v0 = __shfl_sync(0xffffffffu, ln[s][n - 1], lm), w0 = __shfl_sync(0xffffffffu, g1, lm);
          const T v1 = __shfl_sync(0xffffffffu, ln[s][0], lp), w1 = __shfl_sync(0xffffffffu, g0, lp);
          nv[s2][0] = e0 ? hp[s][0].x : v0; nv[s2][1] = e0 ? hp[s][0].y : w0;
          nv[s2][2] = e1 ? hp[s][1].x : v1; nv[s2][3] = e1 ? hp[s][1].y : w1;
        };
        auto stageC = [&](int t, int s, int s2) {
          const bool isrow = t < n;
          const int a = isrow ? t : t - n;
          T out[n];
#pragma unroll
          for (int l = 0; l < n; ++l) out[l] = isrow ? wr[s][l] : xr[l][a];
          if constexpr (EO) line_eo<n, T, true>(Q, isrow ? 0 : 1, ln[s], nv[s2][0], nv[s2][1], nv[s2][2], nv[s2][3], out);
          else {
            const int d = isrow ? 0 : 1;
            line_plain<n, T, true>(P.Kt[d], P.cV[d][0], P.cG[d][0], P.cV[d][1], P.cG[d][1], ln[s], nv[s2][0], nv[s2][1], nv[s2][2], nv[s2][3], out);
          }
          if (isrow) {
#pragma unroll
            for (int l = 0; l < n; ++l) xr[a][l] = out[l];
          } else {
            T res[n];
            line_mass<n, T, EO>(P, Q, 1, out, res);   // mass in y (the Z warps apply Mx and Mz)
#pragma unroll
            for (int l = 0; l < n; ++l) Wp[l * n + a] = res[l];
          }
        };
        stageA(0, 0);
        stageA(1, 1);
        stageB(0, 0, 0);
#pragma unroll
        for (int t = 0; t < 2 * n; ++t) {
          if (t + 2 < 2 * n) stageA(t + 2, (t + 2) % 3);
          if (t + 1 < 2 * n) stageB(t + 1, (t + 1) % 3, (t + 1) & 1);
          stageC(t, t % 3, t & 1);
        }
        // all reads of the stage are done (the W stores above depend on them): the last XY warp refills it
        {
          __syncwarp();
          int last = 0;
          if (lane == 0) {
            const int old = atomicAdd(&ucnt[p % NS], 1);
            if (old == R::NXY - 1) { ucnt[p % NS] = 0; last = 1; }
          }
          last = __shfl_sync(0xffffffffu, last, 0);
          if (last && p + NS < NZ) issue_plane(p + NS);
        }
      } else {
      T u[n][n];
#pragma unroll
      for (int jj = 0; jj < n; ++jj)
#pragma unroll
        for (int i = 0; i < n; ++i) u[jj][i] = Up[jj * n + i];
#pragma unroll
      for (int jj = 0; jj < n; ++jj)
#pragma unroll
        for (int i = 0; i < n; ++i) keep(u[jj][i]);
      // z-partial sums from the Z warps (and the halo traces from the H warps) are ready
      mbar_wait_mode(&full[st], (unsigned)((p >> 1) & 1), G.wait_mode & 1);
      // The layer is in registers: the last XY warp to get here refills the stage with plane p + NS.  Only now, after full(p):
      // the Z warps read plane 0 in their prologue, concurrently with this role (every later plane p is read by the Z warps
      // one iteration ahead, before full(p - 1)); a refill before full(0) would also let ubar[0] run two phases ahead of them.
      {
        __syncwarp();
        int last = 0;
        if (lane == 0) {
          const int old = atomicAdd(&ucnt[p % NS], 1);
          if (old == R::NXY - 1) { ucnt[p % NS] = 0; last = 1; }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && p + NS < NZ) issue_plane(p + NS);
      }
      T acc[n][n];
#pragma unroll
      for (int jj = 0; jj < n; ++jj)
#pragma unroll
        for (int i = 0; i < n; ++i) acc[jj][i] = Wp[jj * n + i];
      // x lines: row jj = face point jj; neighbour traces by shuffle inside the tile, from the halo buffer on the tile edge
#pragma unroll
      for (int jj = 0; jj < n; ++jj) {
        T g0, g1;
        line_traces<n, T, EO>(P, Q, u[jj], g0, g1);   // own outward normal derivatives on the two x faces
        T v0 = __shfl_sync(0xffffffffu, u[jj][n - 1], lxm), w0 = __shfl_sync(0xffffffffu, g1, lxm);
        T v1 = __shfl_sync(0xffffffffu, u[jj][0], lxp), w1 = __shfl_sync(0xffffffffu, g0, lxp);
        { const P2 t = *reinterpret_cast<const P2 *>(hb + hx0 + 2 * jj); v0 = ex0 ? t.x : v0; w0 = ex0 ? t.y : w0; }
        { const P2 t = *reinterpret_cast<const P2 *>(hb + hx1 + 2 * jj); v1 = ex1 ? t.x : v1; w1 = ex1 ? t.y : w1; }
        if constexpr (EO) line_eo<n, T, true>(Q, 0, u[jj], v0, w0, v1, w1, acc[jj]);
        else line_plain<n, T, true>(P.Kt[0], P.cV[0][0], P.cG[0][0], P.cV[0][1], P.cG[0][1], u[jj], v0, w0, v1, w1, acc[jj]);
      }
      // y lines: column i = face point i
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T col[n], out[n], g0, g1;
#pragma unroll
        for (int l = 0; l < n; ++l) { col[l] = u[l][i]; out[l] = acc[l][i]; }
        line_traces<n, T, EO>(P, Q, col, g0, g1);
        T v0 = __shfl_sync(0xffffffffu, col[n - 1], lym), w0 = __shfl_sync(0xffffffffu, g1, lym);
        T v1 = __shfl_sync(0xffffffffu, col[0], lyp), w1 = __shfl_sync(0xffffffffu, g0, lyp);
        { const P2 t = *reinterpret_cast<const P2 *>(hb + hy0 + 2 * i); v0 = ey0 ? t.x : v0; w0 = ey0 ? t.y : w0; }
        { const P2 t = *reinterpret_cast<const P2 *>(hb + hy1 + 2 * i); v1 = ey1 ? t.x : v1; w1 = ey1 ? t.y : w1; }
        if constexpr (EO) line_eo<n, T, true>(Q, 1, col, v0, w0, v1, w1, out);
        else line_plain<n, T, true>(P.Kt[1], P.cV[1][0], P.cG[1][0], P.cV[1][1], P.cG[1][1], col, v0, w0, v1, w1, out);
#pragma unroll
        for (int l = 0; l < n; ++l) acc[l][i] = out[l];
      }
      // mass in y (the Z warps, which hold x-rows and z-columns in registers, apply Mx and Mz)
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T col[n], out[n];
#pragma unroll
        for (int l = 0; l < n; ++l) col[l] = acc[l][i];
        line_mass<n, T, EO>(P, Q, 1, col, out);
#pragma unroll
        for (int l = 0; l < n; ++l) Wp[l * n + i] = out[l];
      }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&qready[st]);
    }
  } else {
    // ------------------------------------------------------------------ H warps: halo traces; warp 0 also issues the bulk stores
    constexpr int NT = R::NH * 32;
    constexpr int SX = (TYC * NL + NT - 1) / NT, SY = (TXC * NL + NT - 1) / NT;   // slots per x side / y side
    constexpr int NSLOT = 2 * SX + 2 * SY;
    const int t = ridx * 32 + lane;
    int hcell[NSLOT], hline[NSLOT], hcn[NSLOT];   // adjacent tile cell (index at plane z0), line offset, halo cell
    auto for_slots = [&](auto &&f) {
      [&]<int... I>(std::integer_sequence<int, I...>) { (f(std::integral_constant<int, I>{}), ...); }(std::make_integer_sequence<int, NSLOT>{});
    };
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
    for (int p = 0; p < NZ + 2; ++p) {
      mbar_wait_mode(&qready[p & 1], (unsigned)(((p - 2) >> 1) & 1), G.wait_mode & 4);   // plane p - 2 is through the XY warps: halo buffer p & 1 is free
      if (p < NZ) {
        // two half batches: NSLOT lines of n values at once would be 120 registers (the first version spilled them)
        T *hb = HXY + (size_t)(p & 1) * HBUF;
        auto half = [&](auto lo_c, auto hi_c) {
          constexpr int LO = decltype(lo_c)::value, HI = decltype(hi_c)::value;
          T x[HI - LO][n];
          auto for_half = [&](auto &&f) {
            [&]<int... I>(std::integer_sequence<int, I...>) { (f(std::integral_constant<int, LO + I>{}), ...); }(std::make_integer_sequence<int, HI - LO>{});
          };
          for_half([&](auto slc) {
            constexpr int sl = decltype(slc)::value;
            constexpr int str = sl < 2 * SX ? 1 : n;
            const T *ptr = src + (size_t)hcn[sl] * ND + hline[sl];
#pragma unroll
            for (int l = 0; l < n; ++l) x[sl - LO][l] = __ldg(ptr + l * str);
          });
          for_half([&](auto slc) {
            constexpr int sl = decltype(slc)::value;
            constexpr int dir = sl < 2 * SX ? 0 : 1, side = dir == 0 ? sl / SX : (sl - 2 * SX) / SY, rnd = dir == 0 ? sl % SX : (sl - 2 * SX) % SY;
            constexpr int per_side = (dir == 0 ? TYC : TXC) * NL;
            constexpr int base = (dir == 0 ? 0 : 2 * TYC * NL) + side * per_side;
            T g0, g1;
            line_traces<n, T, false>(P, Q, x[sl - LO], g0, g1);
            const int e = t + rnd * NT;
            if (e < per_side) {
              P2 v;
              v.x = side ? x[sl - LO][0] : x[sl - LO][n - 1];   // the halo cell's facing side is 1 - side
              v.y = side ? g0 : g1;
              *reinterpret_cast<P2 *>(hb + 2 * (base + e)) = v;
            }
          });
        };
        half(std::integral_constant<int, 0>{}, std::integral_constant<int, NSLOT / 2>{});
        half(std::integral_constant<int, NSLOT / 2>{}, std::integral_constant<int, NSLOT>{});
        const long long sh = plane_term(G, z0 + (p + 1 < NZ ? p + 1 : p)) - pt0;
        for_slots([&](auto slc) {   // halo cells of the next plane
          constexpr int sl = decltype(slc)::value;
          constexpr int dir = sl < 2 * SX ? 0 : 1, side = dir == 0 ? sl / SX : (sl - 2 * SX) / SY;
          hcn[sl] = nbr[(size_t)(2 * dir + side) * no + (hcell[sl] + sh)];
        });
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[p & 1]);
      }
      if (ridx == 0 && p >= 2) {   // bulk-store plane p - 2
        const int ps = p - 2;
        mbar_wait_mode(outfull, (unsigned)(ps & 1), G.wait_mode & 4);
        if (lane < TC / 4) {
          const long long sh = plane_term(G, z0 + ps) - pt0;
          const int ci = cell_index(G, x0 + slot_cx(4 * lane), y0 + slot_cy(4 * lane), z0);
          bulk_s2g(dst + (size_t)(ci + sh) * ND, OUT + (size_t)4 * lane * ND, (unsigned)(4 * ND * sizeof(T)));
          bulk_commit();
          bulk_wait_read0();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(outfree);
      }
    }
    if (ridx == 0 && lane < TC / 4) bulk_wait0();
  }
}

template <typename T, int n>
void fill_eo_params(EOParams<T, n> &Q, const FastParams<double, n> &F) {
  constexpr int NE = (n + 1) / 2, NO = n / 2;
  for (int d = 0; d < 3; ++d) {
    for (int i = 0; i < NE; ++i) {
      for (int l = 0; l < NE; ++l) {
        const bool mid = (n & 1) && l == NE - 1;
        Q.KE[d][i][l] = (T)(mid ? F.Kt[d][i * n + l] : 0.5 * (F.Kt[d][i * n + l] + F.Kt[d][i * n + (n - 1 - l)]));
        Q.ME[d][i][l] = (T)(mid ? F.Mh[d][i * n + l] : 0.5 * (F.Mh[d][i * n + l] + F.Mh[d][i * n + (n - 1 - l)]));
      }
      Q.KEv[d][i] = (T)(0.5 * (F.cV[d][0][i] + F.cV[d][1][i]));
      Q.KEg[d][i] = (T)(0.5 * (F.cG[d][0][i] + F.cG[d][1][i]));
    }
    for (int i = 0; i < NO; ++i) {
      for (int l = 0; l < NO; ++l) {
        Q.KO[d][i][l] = (T)(0.5 * (F.Kt[d][i * n + l] - F.Kt[d][i * n + (n - 1 - l)]));
        Q.MO[d][i][l] = (T)(0.5 * (F.Mh[d][i * n + l] - F.Mh[d][i * n + (n - 1 - l)]));
      }
      Q.KOv[d][i] = (T)(0.5 * (F.cV[d][0][i] - F.cV[d][1][i]));
      Q.KOg[d][i] = (T)(0.5 * (F.cG[d][0][i] - F.cG[d][1][i]));
    }
  }
  // G_0 = sum_l g0_l u_l, G_1 = sum_l g0_{n-1-l} u_l  =>  gs = (G_0 + G_1) / 2, gd = (G_0 - G_1) / 2
  for (int l = 0; l < NE; ++l) {
    const bool mid = (n & 1) && l == NE - 1;
    Q.gE[l] = (T)(mid ? F.g[0][l] : 0.5 * (F.g[0][l] + F.g[0][n - 1 - l]));
  }
  for (int l = 0; l < NO; ++l) Q.gO[l] = (T)(0.5 * (F.g[0][l] - F.g[0][n - 1 - l]));
}

// largest deviation of the 1-D tables from centro-symmetry, relative to their largest entry (the even-odd form needs 0)
template <int n>
double eo_asymmetry(const FastParams<double, n> &F) {
  double dev = 0, mag = 0;
  for (int d = 0; d < 3; ++d) {
    for (int i = 0; i < n; ++i) {
      for (int l = 0; l < n; ++l) {
        dev = std::max(dev, std::fabs(F.Kt[d][i * n + l] - F.Kt[d][(n - 1 - i) * n + (n - 1 - l)]));
        dev = std::max(dev, std::fabs(F.Mh[d][i * n + l] - F.Mh[d][(n - 1 - i) * n + (n - 1 - l)]) * std::fabs(F.Kt[d][0]) / std::fabs(F.Mh[d][0]));
        mag = std::max(mag, std::fabs(F.Kt[d][i * n + l]));
      }
      dev = std::max(dev, std::fabs(F.cV[d][0][i] - F.cV[d][1][n - 1 - i]));
      dev = std::max(dev, std::fabs(F.cG[d][0][i] - F.cG[d][1][n - 1 - i]));
    }
  }
  for (int l = 0; l < n; ++l) dev = std::max(dev, std::fabs(F.g[0][l] - F.g[1][n - 1 - l]) * mag / std::max(1e-300, std::fabs(F.g[0][0])));
  return dev / mag;
}
