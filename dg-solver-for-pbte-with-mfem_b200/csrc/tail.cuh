// End of a source iteration in (at most) two launches, and in ONE on a single GPU:
//
//   iteration_tail_kernel<true>   streams the flux once ([Tc | Qc] partial sums per slot chunk, MacroscopicQuantities::
//                                 AccumulateDirectionalCoeff, src/MacroscopicQuantities.cpp:104-128); the LAST chunk CTA of an
//                                 element block adds the chunk partials in ascending order (fixed order: bit-reproducible) and,
//                                 when no all-reduce has to happen in between (finish = 1), goes straight on to
//   the finish step               cell integrals Tv, Qv (::Finalize :130-157), the three squared norms of the stop test
//                                 (::Residual :159-166 and the two norms the reference prints), the hand-over tc_prev <- Tc,
//                                 tv_prev <- Tv (src/PBTESolver.cpp:328-329); the LAST block of the grid adds the block
//                                 partials in a fixed order, writes {d2, t2, c2, error flag} into the iteration history (pinned
//                                 host memory mapped into the device), advances the iteration counter and raises the STOP flag
//                                 when residual < tol: every kernel of a later iteration returns at once then, so the host may
//                                 enqueue iterations ahead of the residuals it has seen (no synchronisation between iterations)
//                                 and the result is still the reference's: the state of the first iteration with res < tol.
//   iteration_tail_kernel<false>  the same without the streaming pass: sums `nparts` partial vectors that a sweep kernel wrote
//                                 (the per-element kernels write one per upwind-DAG class) or takes the moment vector as it
//                                 is (nparts = 0: after the all-reduce of a multi-GPU iteration), then finishes.
//
// Thread map: a CTA owns EPB = 256 / (np/2) whole elements; a thread owns one 16-byte chunk (two dofs) of one element row, so
// the flux is read with 16-byte loads, 8 slots in flight per thread.
#pragma once

#include "common.cuh"

namespace pbte_dev
{
constexpr int kTailThreads = 256;
constexpr int kTailMaxVals = 512;  // EPB * n <= (256 / ch2) * 2 * ch2

struct DevTail
{
    int nparts;            // STREAM: slot chunks (gridDim.y); else partial vectors to add (0: `mom` already holds the sums)
    int slots_per_chunk;   // STREAM only
    int finish;            // 1: go on to cell integrals / norms / hand-over; 2: cell integrals of the moments [m_begin, m_end) only
    int m_begin, m_end;    // moments the finish part works on (a multi-GPU iteration all-reduces and finishes Tc alone, Qc once at the end)
    double *part;          // [nparts][1+dim][ne*n] partial sums (unused when nparts <= 1 && STREAM, or nparts == 0)
    double *mom;           // [1+dim][ne*n] result of the reduction (the vector a multi-GPU iteration all-reduces)
    unsigned *tickets;     // [gridDim.x + 1] arrival counters (zero between launches)
    const double *basis;   // [ne][n]
    double *cell;          // [1+dim][ne]  Tv | Qv
    double *tv_prev;       // [ne]
    double *tc_prev;       // [ne][np]
    double *block_part;    // [gridDim.x][3]
    DevCtrl *ctrl;
    double *hist;          // device view of the pinned history ring [kHistRing][4]
    // Multi-GPU exchange over NVLink peer memory (instead of an NCCL all-reduce between the two kernels).  Every element block of the
    // tail grid has its own flag, so the exchange of a block starts when that block is ready on every rank, not when the slowest
    // block of the slowest rank is:
    //   peer = 1 (publish): the block's temperature partial sums go into this rank's exchange buffer xchg[parity][ne][np]; the CTA
    //            then stores epoch + 1 into the block's flag (release at system scope);
    //   peer = 2 (finish): waits until every peer has published epoch + 1 for the block, adds the N partial rows in rank order straight
    //            from the peers' memory (P2P loads, identical order on every rank: identical sums), and finishes the iteration;
    //   peer = 3 (marching path: publish, wait, sum and finish in the ONE streaming tail launch): a CTA publishes before it waits and
    //            CTAs start in block order on every rank, so whatever a waiting CTA waits for was started earlier on the peer and
    //            does not wait itself before publishing: no deadlock even when the grid is larger than what is resident.
    // The buffers alternate with the parity of the epoch: a rank can start iteration k + 1 only after every peer has published every
    // block of iteration k, i.e. after every peer is done reading iteration k - 1, whose buffer iteration k + 1 overwrites.
    int peer, nranks, rank;
    double *qpart;                  // peer = 3: where the heat-flux partial sums of this rank go (all-reduced once, at the read-out)
    double *xchg;                   // this rank's buffer: 2 * ne * np doubles, then kXchgFlag0 + number of element blocks flags (long long)
    const double *const *peers;     // [nranks] every rank's buffer as mapped here (cudaIpcOpenMemHandle / peer access)
    long long *epoch;               // iterations exchanged so far (device counter, never reset; advanced by the finish step)
};

constexpr int kXchgFlag0 = 2;  // first per-block flag, in 8-byte units behind the two buffers

__device__ __forceinline__ long long ld_acquire_sys(const long long *p)
{
    long long v;
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(long long *p, long long v)
{
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ bool last_arrival(unsigned *ticket, unsigned expected, int *s_flag)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned t = atomicAdd(ticket, 1u);
        *s_flag = (t == expected - 1) ? 1 : 0;
        if (t == expected - 1) *ticket = 0u;  // ready for the next launch
    }
    __syncthreads();
    const bool last = *s_flag != 0;
    if (last) __threadfence();
    return last;
}

#ifndef PBTE_TAIL_MINBLOCKS
#define PBTE_TAIL_MINBLOCKS 1
#endif
#ifndef PBTE_TAIL_INFLIGHT
#define PBTE_TAIL_INFLIGHT 8
#endif
template <bool STREAM>
__global__ void __launch_bounds__(kTailThreads, PBTE_TAIL_MINBLOCKS) iteration_tail_kernel(DevProblem P, DevTail Z)
{
    if (*P.stop) return;
    __shared__ double sp[kMaxDim + 1][kTailMaxVals];
    __shared__ double sred[3][kTailThreads / 32];
    __shared__ int s_flag;
    const int n = P.n, np = P.np, ne = P.ne, dim = P.dim, nm = 1 + dim;
    const int ch2 = np >> 1, epb = kTailThreads / ch2;
    const int t = threadIdx.x, le = t / ch2, ch = t - le * ch2;
    const int e = blockIdx.x * epb + le;
    const bool valid = le < epb && e < ne;
    const int i0 = 2 * ch;
    const bool has1 = i0 + 1 < n;  // second dof of the chunk exists (n odd: the last chunk holds one dof and the zero padding)
    const size_t block = (size_t)ne * n, x0 = (size_t)e * n + i0;
    double v[kMaxDim + 1][2];
#pragma unroll
    for (int m = 0; m <= kMaxDim; ++m) v[m][0] = v[m][1] = 0.0;

    if (STREAM)
    {
        // slot factors {mf, qf0, qf1, qf2} go through shared memory, 256 slots at a time
        __shared__ double4 s_fac[kTailThreads];
        const int s_begin = blockIdx.y * Z.slots_per_chunk, s_end = min(P.nslots, s_begin + Z.slots_per_chunk);
        const size_t sblock = (size_t)ne * np;
        const double *u = P.flux + (size_t)(valid ? e : 0) * np + i0;
        auto add = [&](const double4 f, const double2 w) {
            v[0][0] = fma(f.x, w.x, v[0][0]);
            v[0][1] = fma(f.x, w.y, v[0][1]);
            v[1][0] = fma(f.y, w.x, v[1][0]);
            v[1][1] = fma(f.y, w.y, v[1][1]);
            v[2][0] = fma(f.z, w.x, v[2][0]);
            v[2][1] = fma(f.z, w.y, v[2][1]);
            v[3][0] = fma(f.w, w.x, v[3][0]);
            v[3][1] = fma(f.w, w.y, v[3][1]);
        };
        for (int sb = s_begin; sb < s_end; sb += kTailThreads)
        {
            const int se = min(s_end, sb + kTailThreads);
            __syncthreads();
            if (sb + t < se) s_fac[t] = make_double4(P.slot_mf[sb + t], P.slot_qf[3 * (sb + t)], P.slot_qf[3 * (sb + t) + 1], P.slot_qf[3 * (sb + t) + 2]);
            __syncthreads();
            if (!valid) continue;
            int s = sb;
            constexpr int NF = PBTE_TAIL_INFLIGHT;  // 16-byte loads in flight per thread
            for (; s + NF <= se; s += NF)
            {
                double2 w[NF];
#pragma unroll
                for (int k = 0; k < NF; ++k) w[k] = __ldcs(reinterpret_cast<const double2 *>(u + (size_t)(s + k) * sblock));
#pragma unroll
                for (int k = 0; k < NF; ++k) add(s_fac[s + k - sb], w[k]);
            }
            for (; s < se; ++s) add(s_fac[s - sb], __ldcs(reinterpret_cast<const double2 *>(u + (size_t)s * sblock)));
        }
        if (Z.nparts > 1)
        {
            if (valid)
            {
                double *pm = Z.part + (size_t)blockIdx.y * nm * block + x0;
#pragma unroll
                for (int m = 0; m <= kMaxDim; ++m)
                {
                    if (m >= nm) break;
                    pm[(size_t)m * block] = v[m][0];
                    if (has1) pm[(size_t)m * block + 1] = v[m][1];
                }
            }
            if (!last_arrival(Z.tickets + blockIdx.x, (unsigned)Z.nparts, &s_flag)) return;
        }
    }
    // ---- multi-GPU: wait for the peers' partial sums of this iteration ------------------------------------------------------------
    const size_t xrow = (size_t)ne * np;  // doubles per exchange buffer
    long long ep = 0;
    auto wait_for_peers = [&]() {
        if (t < Z.nranks && t != Z.rank)
        {
            const long long *flag = reinterpret_cast<const long long *>(Z.peers[t] + 2 * xrow) + kXchgFlag0 + blockIdx.x;
            const long long t0 = clock64();
            while (ld_acquire_sys(flag) < ep + 1)
            {
                __nanosleep(100);
                if (clock64() - t0 > 20000000000LL)
                {
                    // a peer that never arrives (its process died): give up loudly instead of hanging the GPU
                    atomicCAS(P.error_flag, 0, PBTE_ERR_CUDA);
                    break;
                }
            }
        }
        __syncthreads();
    };
    // the N ranks' partial sums of this thread's chunk in rank order, straight from their memory (uncached: the buffers are rewritten
    // every other iteration)
    auto sum_over_ranks = [&](double &a, double &b) {
        const size_t off = (size_t)(ep & 1) * xrow + (size_t)e * np + i0;
        a = 0.0;
        b = 0.0;
        for (int rk = 0; rk < Z.nranks; ++rk)
        {
            const double2 w = __ldcv(reinterpret_cast<const double2 *>(Z.peers[rk] + off));
            a += w.x;
            b += w.y;
        }
    };
    if (Z.peer)
    {
        ep = *Z.epoch;
        if (Z.peer == 2) wait_for_peers();
    }
    // ---- sum of the partial vectors, ascending (fixed order) ---------------------------------------------------------------
    if (valid)
    {
        const int nparts = (STREAM && Z.nparts <= 1) ? -1 : Z.nparts;  // -1: v already holds the sums
        if (nparts >= 0)
        {
#pragma unroll
            for (int m = 0; m <= kMaxDim; ++m)
            {
                if (m >= nm) break;
                if (nparts == 0 && (m < Z.m_begin || m >= Z.m_end)) continue;
                double a = 0.0, b = 0.0;
                if (nparts == 0 && Z.peer == 2 && m == 0)
                {
                    // the all-reduced row also goes to the local moment vector
                    sum_over_ranks(a, b);
                    Z.mom[x0] = a;
                    if (has1) Z.mom[x0 + 1] = b;
                }
                else if (nparts == 0)
                {
                    a = Z.mom[(size_t)m * block + x0];
                    if (has1) b = Z.mom[(size_t)m * block + x0 + 1];
                }
                for (int y = 0; y < nparts; ++y)
                {
                    const double *pm = Z.part + ((size_t)y * nm + m) * block + x0;
                    a += __ldcg(pm);
                    if (has1) b += __ldcg(pm + 1);
                }
                v[m][0] = a;
                v[m][1] = b;
            }
        }
        if (nparts != 0)
#pragma unroll
            for (int m = 0; m <= kMaxDim; ++m)
            {
                if (m >= nm) break;
                if (m == 0 && (Z.peer == 1 || Z.peer == 3))
                {
                    // temperature partial sums of this rank: into the exchange buffer the peers read (padded rows: 16-byte stores)
                    *reinterpret_cast<double2 *>(Z.xchg + (size_t)(ep & 1) * xrow + (size_t)e * np + i0) = make_double2(v[0][0], has1 ? v[0][1] : 0.0);
                    continue;
                }
                double *dst = Z.peer == 3 ? Z.qpart : Z.mom;
                dst[(size_t)m * block + x0] = v[m][0];
                if (has1) dst[(size_t)m * block + x0 + 1] = v[m][1];
            }
    }
    if (Z.peer == 1 || Z.peer == 3)
    {
        // publish the block: every thread's rows are written (CTA barrier), then one release at system scope
        __syncthreads();
        if (t == 0)
        {
            __threadfence_system();
            st_release_sys(reinterpret_cast<long long *>(Z.xchg + 2 * xrow) + kXchgFlag0 + blockIdx.x, ep + 1);
        }
    }
    if (Z.peer == 3)
    {
        wait_for_peers();
        if (valid)
        {
            sum_over_ranks(v[0][0], v[0][1]);
            Z.mom[x0] = v[0][0];
            if (has1) Z.mom[x0 + 1] = v[0][1];
        }
    }
    if (!Z.finish) return;

    // ---- finish: cell integrals, norms, hand-over ----------------------------------------------------------------------------
    const int mb = Z.m_begin, me = min(Z.m_end, nm);
    double c2 = 0.0;
    if (valid)
    {
        const double b0 = Z.basis[x0], b1 = has1 ? Z.basis[x0 + 1] : 0.0;
#pragma unroll
        for (int m = 0; m <= kMaxDim; ++m)
        {
            if (m >= me) break;
            if (m < mb) continue;
            sp[m][le * n + i0] = v[m][0] * b0;
            if (has1) sp[m][le * n + i0 + 1] = v[m][1] * b1;
        }
        if (mb == 0 && Z.finish == 1)
        {
            // tc_prev rows are padded to np doubles; the padding stays zero
            *reinterpret_cast<double2 *>(Z.tc_prev + (size_t)e * np + i0) = make_double2(v[0][0], has1 ? v[0][1] : 0.0);
            c2 = v[0][0] * v[0][0] + (has1 ? v[0][1] * v[0][1] : 0.0);
        }
    }
    __syncthreads();
    double d2 = 0.0, t2 = 0.0;
    {
        // thread (m, element): sum over the dofs in ascending order, as the reference's loop does
        const int nel = min(epb, ne - (int)blockIdx.x * epb);
        for (int idx = t; idx < (me - mb) * nel; idx += kTailThreads)
        {
            const int m = mb + idx / nel, l = idx - (m - mb) * nel, ee = blockIdx.x * epb + l;
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += sp[m][l * n + i];
            Z.cell[(size_t)m * ne + ee] = s;
            if (m == 0 && Z.finish == 1)
            {
                const double d = s - Z.tv_prev[ee];
                d2 += d * d;
                t2 += s * s;
                Z.tv_prev[ee] = s;
            }
        }
    }
    if (Z.finish == 2) return;
    // CTA partial sums in a fixed order: warp shuffle tree, then the warps in ascending order
    for (int off = 16; off > 0; off >>= 1)
    {
        d2 += __shfl_down_sync(0xffffffffu, d2, off);
        t2 += __shfl_down_sync(0xffffffffu, t2, off);
        c2 += __shfl_down_sync(0xffffffffu, c2, off);
    }
    const int lane = t & 31, warp = t >> 5;
    if (lane == 0)
    {
        sred[0][warp] = d2;
        sred[1][warp] = t2;
        sred[2][warp] = c2;
    }
    __syncthreads();
    if (t < 3)
    {
        double a = 0.0;
        for (int w = 0; w < kTailThreads / 32; ++w) a += sred[t][w];
        Z.block_part[(size_t)blockIdx.x * 3 + t] = a;
    }
    if (!last_arrival(Z.tickets + gridDim.x, gridDim.x, &s_flag)) return;
    // ---- last block of the grid: the three norms, history entry, stop test ----------------------------------------------------
    double a[3] = {0.0, 0.0, 0.0};
    for (int b = t; b < (int)gridDim.x; b += kTailThreads)
#pragma unroll
        for (int k = 0; k < 3; ++k) a[k] += __ldcg(Z.block_part + (size_t)b * 3 + k);
#pragma unroll
    for (int k = 0; k < 3; ++k)
        for (int off = 16; off > 0; off >>= 1) a[k] += __shfl_down_sync(0xffffffffu, a[k], off);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) sred[k][warp] = a[k];
    __syncthreads();
    if (t == 0)
    {
        double tot[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            double s = 0.0;
            for (int w = 0; w < kTailThreads / 32; ++w) s += sred[k][w];
            tot[k] = s;
        }
        DevCtrl *c = Z.ctrl;
        const int it = c->iter;
        double *h = Z.hist + (size_t)(it % kHistRing) * 4;
        h[0] = tot[0];
        h[1] = tot[1];
        h[2] = tot[2];
        h[3] = (double)*P.error_flag;
        c->iter = it + 1;
        if (Z.peer >= 2) *Z.epoch = ep + 1;
        // the reference's stop test: res = ||Tv - Tv_prev|| / ||Tv|| < tol (src/PBTESolver.cpp:327)
        if (tot[1] > 0.0 && sqrt(tot[0]) / sqrt(tot[1]) < c->tol) c->stop = 1;
        __threadfence_system();
    }
}

}  // namespace pbte_dev
