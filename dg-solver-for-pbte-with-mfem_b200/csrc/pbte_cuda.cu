// libpbte_cuda: C ABI of include/pbte_cuda.h.  Host-side orchestration (device image, schedule classes,
// level lists, CUDA-graph replay of a sweep) around the kernels in schedule.cuh / sweep_*.cuh / moments.cuh.
#include "pbte_cuda.h"

#include "common.cuh"
#include "moments.cuh"
#include "peaks.cuh"
#include "schedule.cuh"
#include "shared_plan.hpp"
#include "sweep_generic.cuh"
#include "sweep_regop.cuh"
#include "sweep_shared.cuh"
#include "sweep_warp.cuh"
#include "tail.cuh"

#include <nccl.h>
#include <nvtx3/nvToolsExt.h>  // header-only: ranges around the library's phases for nsys / ncu timelines (tracing)

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace pbte_dev;

namespace
{
struct NvtxRange
{
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

thread_local std::string g_err = "no error";

int fail(int status, const std::string &msg)
{
    g_err = msg;
    return status;
}

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t ck_e_ = (call);                                                                           \
        if (ck_e_ != cudaSuccess)                                                                             \
            return fail(PBTE_ERR_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(ck_e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

template <class T>
struct DevBuf
{
    T *p = nullptr;
    size_t count = 0;
    cudaError_t alloc(size_t n)
    {
        release();
        count = n;
        if (n == 0) return cudaSuccess;
        return cudaMalloc(&p, n * sizeof(T));
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        count = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};
}  // namespace

struct pbte_handle
{
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int cache_policy = 0, kernel_choice = 0, use_graph = 1;

    // sizes
    int dim = 0, ne = 0, n = 0, np = 0, F = 0, ndir = 0, nband = 0, ms = 0;  // np: padded row stride of flux / tc_prev
    int k0 = 0, k1 = 0;  // owned directions [k0, k1)
    int nloc() const { return k1 - k0; }
    int nslots() const { return nloc() * nband; }
    bool has_topology = false, has_matrices = false, has_bands = false, has_basis = false, has_dirs = false, missing_bc = false;
    double omega = 0.0, heat_cap_v = 0.0, nu_max = 0.0;
    std::vector<double> h_dir, h_dirw, h_invkn, h_vg, h_hc, h_fw;
    std::vector<int32_t> h_fnbr;
    std::vector<double> h_fnormal;

    // device image
    DevBuf<double> mass, massT, stiff, fmass, fcoup, fint, fnormal, ftbc, basis, dirs;
    DevBuf<int32_t> fnbr;

    // schedule
    bool scheduled = false;
    int nclass = 0, max_levels = 0;
    std::vector<int> class_of_dir;         // [nloc]
    std::vector<int> slotdir_of_dir;       // [nloc] position of the direction in class-major order
    std::vector<int> cls_slot_begin;       // [nclass+1] (in slots)
    std::vector<int32_t> h_pass, h_level;  // [nclass][ne]
    std::vector<int> level_offsets;        // [max_levels+1] into items
    DevBuf<int32_t> item_done;             // [nclass][ne] "stored" flags of the whole-sweep launch of the per-element kernel
    int warp_grid = 0;                     // its grid: what is resident at once (0: one launch per wavefront level)
    DevBuf<int2> items;
    DevBuf<int32_t> d_cls_slot_begin;
    DevBuf<uint8_t> cls_out, cls_in, pivots;
    DevBuf<double> slot_dir, slot_v, slot_c1, slot_c2, slot_cb, slot_mf, slot_qf;
    WarpKernelShape shape{0, 0};
    int kernel_used = 0;
    bool pivots_recorded = false;

    // shared-operator (marching) path
    double class_tol = 1e-12;
    int max_classes = 512;
    pbte_host::SharedPlan plan;
    int march_ntiles = 0;
    int64_t march_mask_mismatch = 0;
    double march_fill = 0.0;
    bool march_stage = false;
    DevBuf<long long> m_trace;  // development builds only
    bool march_regop = false;  // one operator class: operators in consumer registers (sweep_regop.cuh)
    int march_h_smem = 0;
    int march_ring = 0, march_nv = 0;  // tile-ring slots per CTA, staged vectors per slot
    int march_op_layout = 0;           // DevMarch::op_layout: 0 fragment pairs, 1 compact row-major, 2 split order
    int march_op_stride = 0;           // doubles per operator in m_ops
    int march_max_op_cnt = 0, march_max_class_tiles = 0;
    bool march_persist = false;             // sweep_march_kernel (persistent CTAs, work lists) instead of the one-CTA-per-slot classic kernel
    int march_ncta = 0, march_npieces = 0;  // persistent CTAs of sweep_march_kernel and the pieces of their work lists
    DevBuf<int32_t> m_cta_piece_begin, m_pieces, m_progress;
    size_t march_smem = 0;
    DevBuf<int32_t> m_slot_class, m_cls_level_begin, m_level_tile_begin, m_tile_oclass, m_tile_desc, m_cls_op_off, m_cls_op_cnt;
    DevBuf<uint8_t> m_cls_inmask;
    // local matrices of the operator-class representatives (all the marching path needs of the matrix image)
    DevBuf<double> r_mass, r_massT, r_stiff, r_fmass, r_fcoup, r_fint, r_normal;
    DevBuf<uint8_t> r_has_k, r_has_g, r_cface, r_present;
    bool reps_only = false;      // the per-element matrix image was not uploaded: marching path only
    int64_t image_bytes = 0;     // device bytes of local matrices held (per-element image and/or representatives)
    double create_ms = 0.0;
    DevBuf<unsigned long long> m_slot_op_base;
    DevBuf<double> m_ops, m_oph, m_ctbc;

    // state
    DevBuf<double> flux, tc_prev, tv_prev, partial, mom, mom_part, cell, norms, scratch, block_part;
    DevBuf<int32_t> error_flag;
    double *h_norms = nullptr;  // pinned [4]
    int32_t *h_flag = nullptr;  // pinned
    double last_tv_norm = 0.0, last_tc_norm = 0.0;

    // iteration control (csrc/tail.cuh): device-side stop test, history ring in pinned host memory mapped into the device
    DevBuf<DevCtrl> ctrl;
    DevBuf<unsigned> tickets;
    double *h_hist = nullptr, *d_hist = nullptr;  // [kHistRing][4] = d2, t2, c2, error flag of every finished iteration
    int64_t iter_enq = 0;                         // finish steps enqueued since the device counter was last reset
    int tail_blocks = 0, tail_chunks = 1;         // grid of the tail kernels
    static constexpr int kLookMax = 16;
    cudaEvent_t ev_iter[kLookMax] = {};           // one event per in-flight iteration of pbte_cuda_solve

    // bulk transfer of all coefficient blocks (pbte_cuda_upload_coeff_all / download_coeff_all)
    cudaStream_t xfer = nullptr;
    cudaEvent_t ev_ready[2] = {}, ev_free[2] = {};
    DevBuf<double> stage;               // two staging buffers of stage_blocks blocks each
    DevBuf<int32_t> slot_of_block;      // [nloc * nband] flux slot of host block (k - k0) * nband + band
    int stage_blocks = 0;

    // communication: one all-reduce of [Tc | Qc] per iteration, issued by the library when a communicator is attached
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    bool qc_pending = false;  // the heat-flux moments of the last iteration are not all-reduced yet (complete_moments)
    // exchange of the temperature partial sums over NVLink peer memory (csrc/tail.cuh): this rank's buffer, the peers' buffers as
    // mapped into this process, the epoch counter
    bool peer_mode = false;
    bool peer_fused = true;        // marching path: publish, wait, sum and finish inside the one tail launch (PBTE_PEER_FUSED=0: two launches)
    bool sweep_finished = false;   // set by enqueue_sweep: the tail it enqueued has done the exchange and the finish step
    DevBuf<double> xchg;
    DevBuf<long long> xchg_epoch;
    DevBuf<const double *> d_peers;
    std::vector<void *> ipc_opened;  // peer mappings opened with cudaIpcOpenMemHandle (closed at detach)

    // graph replay: [0] sweep + partial moments (the caller or NCCL reduces), [1] sweep + fused finish (whole single-GPU iteration),
    // [2] sweep + NCCL all-reduce + finish (whole multi-GPU iteration of this rank)
    cudaGraphExec_t graph_exec[3] = {nullptr, nullptr, nullptr};
    int graph_mode[3] = {-1, -1, -1};
    bool graph_failed[3] = {false, false, false};  // capture or instantiation of this kind of graph failed once: plain launches from then on
    bool nccl_in_graph = true;

    // stats
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    bool sweep_timed = false;
    int64_t launches = 0, sweeps = 0, launches_per_sweep = 0;
    double last_sweep_ms = 0.0, last_finish_ms = 0.0, last_march_ms = 0.0;
    cudaEvent_t ev_march = nullptr;  // recorded right after the marching kernel (an event-record node when the sweep is a graph)

    void drop_graphs()
    {
        for (int g = 0; g < 3; ++g)
        {
            if (graph_exec[g]) cudaGraphExecDestroy(graph_exec[g]);
            graph_exec[g] = nullptr;
            graph_mode[g] = -1;
        }
    }

    ~pbte_handle()
    {
        drop_graphs();
        for (void *q : ipc_opened) cudaIpcCloseMemHandle(q);
        if (comm) ncclCommDestroy(comm);
        for (cudaEvent_t e : ev_iter)
            if (e) cudaEventDestroy(e);
        if (h_hist) cudaFreeHost(h_hist);
        for (int b = 0; b < 2; ++b)
        {
            if (ev_ready[b]) cudaEventDestroy(ev_ready[b]);
            if (ev_free[b]) cudaEventDestroy(ev_free[b]);
        }
        if (xfer) cudaStreamDestroy(xfer);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev2) cudaEventDestroy(ev2);
        if (ev_march) cudaEventDestroy(ev_march);
        if (h_norms) cudaFreeHost(h_norms);
        if (h_flag) cudaFreeHost(h_flag);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

namespace
{
struct DeviceGuard
{
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// upload `count` matrices of n*n doubles with a device stride of ms doubles
int upload_matrices(DevBuf<double> &dst, const double *src, size_t count, int n, int ms, cudaStream_t st)
{
    const size_t nn = (size_t)n * n;
    CK(dst.alloc(count * ms));
    if (count == 0) return PBTE_OK;
    if ((size_t)ms == nn) CK(cudaMemcpyAsync(dst.p, src, count * nn * sizeof(double), cudaMemcpyHostToDevice, st));
    else
    {
        CK(cudaMemsetAsync(dst.p, 0, count * ms * sizeof(double), st));
        CK(cudaMemcpy2DAsync(dst.p, ms * sizeof(double), src, nn * sizeof(double), nn * sizeof(double), count, cudaMemcpyHostToDevice, st));
    }
    return PBTE_OK;
}

template <class T>
int upload(DevBuf<T> &dst, const T *src, size_t count, cudaStream_t st)
{
    CK(dst.alloc(count));
    if (count) CK(cudaMemcpyAsync(dst.p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    return PBTE_OK;
}

template <class T>
int upload_vec(DevBuf<T> &dst, const T *src, size_t count, cudaStream_t st)
{
    return upload(dst, src, count, st);
}

DevProblem make_dev_problem(pbte_handle *h)
{
    DevProblem P{};
    P.dim = h->dim;
    P.ne = h->ne;
    P.n = h->n;
    P.np = h->np;
    P.F = h->F;
    P.nband = h->nband;
    P.nslots = h->nslots();
    P.ms = h->ms;
    P.nclass = h->nclass;
    P.nu_max = h->nu_max;
    P.mass = h->mass.p;
    P.massT = h->massT.p;
    P.stiff = h->stiff.p;
    P.fmass = h->fmass.p;
    P.fcoup = h->fcoup.p;
    P.fint = h->fint.p;
    P.fnormal = h->fnormal.p;
    P.ftbc = h->ftbc.p;
    P.basis = h->basis.p;
    P.fnbr = h->fnbr.p;
    P.slot_dir = h->slot_dir.p;
    P.slot_v = h->slot_v.p;
    P.slot_c1 = h->slot_c1.p;
    P.slot_c2 = h->slot_c2.p;
    P.slot_cb = h->slot_cb.p;
    P.slot_mf = h->slot_mf.p;
    P.slot_qf = h->slot_qf.p;
    P.flux = h->flux.p;
    P.tc_prev = h->tc_prev.p;
    P.partial = h->partial.p;
    P.cls_slot_begin = h->d_cls_slot_begin.p;
    P.cls_out = h->cls_out.p;
    P.cls_in = h->cls_in.p;
    P.pivots = h->pivots.p;
    P.error_flag = h->error_flag.p;
    P.stop = h->ctrl.p ? &h->ctrl.p->stop : nullptr;
    return P;
}

// ---- warp-kernel dispatch ------------------------------------------------------------------------
typedef void (*WarpKernelFn)(DevProblem, const int2 *, int, int, int32_t *, WarpKernelShape);

template <int N, int DIM>
WarpKernelFn pick_piv(int piv)
{
    switch (piv)
    {
        case 0: return sweep_warp_kernel<N, DIM, 0>;
        case 1: return sweep_warp_kernel<N, DIM, 1>;
        default: return sweep_warp_kernel<N, DIM, 2>;
    }
}

WarpKernelFn pick_warp_kernel(int n, int dim, int piv)
{
    if (dim == 2)
    {
        switch (n)
        {
            case 3: return pick_piv<3, 2>(piv);
            case 4: return pick_piv<4, 2>(piv);
            case 6: return pick_piv<6, 2>(piv);
            case 9: return pick_piv<9, 2>(piv);
            case 10: return pick_piv<10, 2>(piv);
            case 15: return pick_piv<15, 2>(piv);
            case 16: return pick_piv<16, 2>(piv);
            case 25: return pick_piv<25, 2>(piv);
            default: return nullptr;
        }
    }
    if (dim == 3)
    {
        switch (n)
        {
            case 4: return pick_piv<4, 3>(piv);
            case 8: return pick_piv<8, 3>(piv);
            case 10: return pick_piv<10, 3>(piv);
            case 20: return pick_piv<20, 3>(piv);
            case 27: return pick_piv<27, 3>(piv);
            default: return nullptr;
        }
    }
    return nullptr;
}

DevMarch make_dev_march(pbte_handle *h)
{
    DevMarch S{};
    S.noclass = h->plan.noclass;
    S.h_smem = h->march_h_smem;
    S.ring = h->march_ring;
    S.op_stride = h->march_op_stride;
    S.op_layout = h->march_op_layout;
    S.max_op_cnt = h->march_max_op_cnt;
    S.max_class_tiles = h->march_max_class_tiles;
    S.ncta = h->march_ncta;
    S.cta_piece_begin = h->m_cta_piece_begin.p;
    S.pieces = h->m_pieces.p;
    S.progress = h->m_progress.p;
    S.nv = h->march_nv;
    S.F = h->F;
    S.slot_class = h->m_slot_class.p;
    S.cls_level_begin = h->m_cls_level_begin.p;
    S.level_tile_begin = h->m_level_tile_begin.p;
    S.tile_oclass = h->m_tile_oclass.p;
    S.tile_desc = h->m_tile_desc.p;
    S.cls_inmask = h->m_cls_inmask.p;
    S.cls_op_off = h->m_cls_op_off.p;
    S.cls_op_cnt = h->m_cls_op_cnt.p;
    S.slot_op_base = h->m_slot_op_base.p;
    S.ops = h->m_ops.p;
    S.oph = h->m_oph.p;
    S.ctbc = h->m_ctbc.p;
    S.rep_mass = h->r_mass.p;
    S.rep_massT = h->r_massT.p;
    S.rep_stiff = h->r_stiff.p;
    S.rep_fmass = h->r_fmass.p;
    S.rep_fcoup = h->r_fcoup.p;
    S.rep_fint = h->r_fint.p;
    S.rep_normal = h->r_normal.p;
    S.rep_has_k = h->r_has_k.p;
    S.rep_has_g = h->r_has_g.p;
    S.rep_cface = h->r_cface.p;
    S.rep_present = h->r_present.p;
    S.trace = h->m_trace.p;
    return S;
}

// ---- marching-kernel dispatch --------------------------------------------------------------------
typedef void (*MarchKernelFn)(DevProblem, DevMarch);

template <int N>
MarchKernelFn pick_stage(bool stage)
{
    return stage ? sweep_march_kernel<N, true> : sweep_march_kernel<N, false>;
}

MarchKernelFn pick_march_kernel(int n, bool stage)
{
    switch (n)
    {
        case 3: return pick_stage<3>(stage);
        case 4: return pick_stage<4>(stage);
        case 6: return pick_stage<6>(stage);
        case 8: return pick_stage<8>(stage);
        case 9: return pick_stage<9>(stage);
        case 10: return pick_stage<10>(stage);
        case 15: return pick_stage<15>(stage);
        case 16: return pick_stage<16>(stage);
        case 20: return pick_stage<20>(stage);
        case 25: return pick_stage<25>(stage);
        case 27: return pick_stage<27>(stage);
        case 35: return pick_stage<35>(stage);
        case 36: return pick_stage<36>(stage);
        case 64: return pick_stage<64>(stage);
        default: return nullptr;
    }
}

template <int N>
MarchKernelFn pick_classic_stage(bool stage)
{
    return stage ? sweep_march_classic_kernel<N, true> : sweep_march_classic_kernel<N, false>;
}

MarchKernelFn pick_classic_kernel(int n, bool stage)
{
    switch (n)
    {
        case 3: return pick_classic_stage<3>(stage);
        case 4: return pick_classic_stage<4>(stage);
        case 6: return pick_classic_stage<6>(stage);
        case 8: return pick_classic_stage<8>(stage);
        case 9: return pick_classic_stage<9>(stage);
        case 10: return pick_classic_stage<10>(stage);
        case 15: return pick_classic_stage<15>(stage);
        case 16: return pick_classic_stage<16>(stage);
        case 20: return pick_classic_stage<20>(stage);
        case 25: return pick_classic_stage<25>(stage);
        case 27: return pick_classic_stage<27>(stage);
        case 35: return pick_classic_stage<35>(stage);
        case 36: return pick_classic_stage<36>(stage);
        case 64: return pick_classic_stage<64>(stage);
        default: return nullptr;
    }
}

template <int N>
MarchKernelFn pick_regop_n()
{
    if constexpr (regop_supported(N)) return sweep_march_regop_kernel<N>;
    return nullptr;
}

MarchKernelFn pick_regop_kernel(int n)
{
    switch (n)
    {
        case 4: return pick_regop_n<4>();
        case 8: return pick_regop_n<8>();
        case 9: return pick_regop_n<9>();
        case 16: return pick_regop_n<16>();
        case 25: return pick_regop_n<25>();
        case 27: return pick_regop_n<27>();
        case 36: return pick_regop_n<36>();
        case 64: return pick_regop_n<64>();
        default: return nullptr;
    }
}

// Local matrices of the operator-class representatives, gathered from the caller's arrays (tens of elements at most): all that the
// operator construction of the marching path reads (sweep_shared.cuh: build_operators_kernel).
int upload_representatives(pbte_handle *h, const pbte_problem_t *p)
{
    const pbte_host::SharedPlan &plan = h->plan;
    const int noc = plan.noclass, n = h->n, F = h->F, dim = h->dim;
    const size_t nn = (size_t)n * n, ms = (size_t)h->ms;
    std::vector<double> mass(noc * ms, 0.0), massT(noc * ms, 0.0), stiff((size_t)noc * dim * ms, 0.0), fmass((size_t)noc * F * ms, 0.0),
        fcoup((size_t)noc * F * ms, 0.0), fint((size_t)noc * F * n, 0.0), normal((size_t)noc * F * 3, 0.0);
    std::vector<uint8_t> has_k((size_t)noc * F, 0), has_g((size_t)noc * F, 0), cface((size_t)noc * F, 0), present((size_t)noc * F, 0);
    for (int o = 0; o < noc; ++o)
    {
        const size_t e = (size_t)plan.oc_rep[o];
        std::copy(p->mass + e * nn, p->mass + (e + 1) * nn, mass.begin() + o * ms);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) massT[o * ms + i + (size_t)j * n] = p->mass[e * nn + j + (size_t)i * n];
        for (int d = 0; d < dim; ++d) std::copy(p->stiff + (e * dim + d) * nn, p->stiff + (e * dim + d + 1) * nn, stiff.begin() + ((size_t)o * dim + d) * ms);
        for (int f = 0; f < F; ++f)
        {
            std::copy(p->face_mass + (e * F + f) * nn, p->face_mass + (e * F + f + 1) * nn, fmass.begin() + ((size_t)o * F + f) * ms);
            std::copy(p->face_normal + (e * F + f) * 3, p->face_normal + (e * F + f + 1) * 3, normal.begin() + ((size_t)o * F + f) * 3);
        }
        for (int j = 0; j < F; ++j)
        {
            const size_t oj = (size_t)o * F + j;
            cface[oj] = plan.cface[e * F + j];
            present[oj] = plan.cnbr[e * F + j] != PBTE_FACE_NONE;
            const int ks = plan.oc_ksrc[oj], gs = plan.oc_gsrc[oj];
            if (ks >= 0)
            {
                has_k[oj] = 1;
                std::copy(p->face_coupling + (size_t)ks * nn, p->face_coupling + ((size_t)ks + 1) * nn, fcoup.begin() + oj * ms);
            }
            if (gs >= 0)
            {
                has_g[oj] = 1;
                std::copy(p->face_integral + (size_t)gs * n, p->face_integral + ((size_t)gs + 1) * n, fint.begin() + oj * n);
            }
        }
    }
    int st;
    if ((st = upload(h->r_mass, mass.data(), mass.size(), h->stream))) return st;
    if ((st = upload(h->r_massT, massT.data(), massT.size(), h->stream))) return st;
    if ((st = upload(h->r_stiff, stiff.data(), stiff.size(), h->stream))) return st;
    if ((st = upload(h->r_fmass, fmass.data(), fmass.size(), h->stream))) return st;
    if ((st = upload(h->r_fcoup, fcoup.data(), fcoup.size(), h->stream))) return st;
    if ((st = upload(h->r_fint, fint.data(), fint.size(), h->stream))) return st;
    if ((st = upload(h->r_normal, normal.data(), normal.size(), h->stream))) return st;
    if ((st = upload(h->r_has_k, has_k.data(), has_k.size(), h->stream))) return st;
    if ((st = upload(h->r_has_g, has_g.data(), has_g.size(), h->stream))) return st;
    if ((st = upload(h->r_cface, cface.data(), cface.size(), h->stream))) return st;
    if ((st = upload(h->r_present, present.data(), present.size(), h->stream))) return st;
    CK(cudaStreamSynchronize(h->stream));  // the host vectors go out of scope
    h->image_bytes += (int64_t)((mass.size() + massT.size() + stiff.size() + fmass.size() + fcoup.size() + fint.size() + normal.size()) * sizeof(double));
    return PBTE_OK;
}

// Coefficient blocks between the caller's layout (block j = (k - k0) * nband + band: ndof x ne column-major, contiguous) in a
// staging buffer and the flux (slot-major, rows padded to np).  One thread per padded entry.
template <bool UPLOAD>
__global__ void __launch_bounds__(256) repack_blocks_kernel(int nblocks, int ne, int n, int np, const int32_t *__restrict__ slot_of_block, int j0,
                                                           double *__restrict__ stage, double *__restrict__ flux)
{
    const size_t per = (size_t)ne * np, idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)nblocks * per) return;
    const int j = (int)(idx / per);
    const size_t r = idx - (size_t)j * per, e = r / np;
    const int i = (int)(r - e * np);
    double *f = flux + (size_t)slot_of_block[j0 + j] * per + r;
    double *sg = stage + ((size_t)j * ne + e) * n + i;
    if (UPLOAD) *f = i < n ? *sg : 0.0;
    else if (i < n) *sg = *f;
}

int transfer_all_blocks(pbte_handle *h, double *host, bool upload)
{
    NvtxRange nvtx(upload ? "pbte_cuda_upload_coeff_all" : "pbte_cuda_download_coeff_all");
    const int nblk = h->nslots();
    if (nblk == 0) return PBTE_OK;
    const size_t block = (size_t)h->ne * h->n;
    if (!h->xfer)
    {
        CK(cudaStreamCreateWithFlags(&h->xfer, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b)
        {
            CK(cudaEventCreateWithFlags(&h->ev_ready[b], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->ev_free[b], cudaEventDisableTiming));
        }
    }
    if (h->slot_of_block.count != (size_t)nblk)
    {
        std::vector<int32_t> tbl(nblk);
        for (int k = 0; k < h->nloc(); ++k)
            for (int b = 0; b < h->nband; ++b) tbl[(size_t)k * h->nband + b] = h->slotdir_of_dir[k] * h->nband + b;
        const int st = upload_vec(h->slot_of_block, tbl.data(), tbl.size(), h->stream);
        if (st) return st;
        CK(cudaStreamSynchronize(h->stream));
    }
    if (h->stage_blocks == 0)
    {
        h->stage_blocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)nblk, ((size_t)32 << 20) / (block * sizeof(double))));
        CK(h->stage.alloc(2 * (size_t)h->stage_blocks * block));
    }
    const int per = h->stage_blocks;
    CK(cudaStreamSynchronize(h->stream));
    int c = 0;
    for (int j0 = 0; j0 < nblk; j0 += per, ++c)
    {
        const int b = c & 1, cnt = std::min(per, nblk - j0);
        double *sg = h->stage.p + (size_t)b * per * block, *hp = host + (size_t)j0 * block;
        const size_t bytes = (size_t)cnt * block * sizeof(double);
        const unsigned grid = (unsigned)(((size_t)cnt * h->ne * h->np + 255) / 256);
        if (upload)
        {
            if (c >= 2) CK(cudaStreamWaitEvent(h->xfer, h->ev_free[b], 0));
            CK(cudaMemcpyAsync(sg, hp, bytes, cudaMemcpyHostToDevice, h->xfer));
            CK(cudaEventRecord(h->ev_ready[b], h->xfer));
            CK(cudaStreamWaitEvent(h->stream, h->ev_ready[b], 0));
            repack_blocks_kernel<true><<<grid, 256, 0, h->stream>>>(cnt, h->ne, h->n, h->np, h->slot_of_block.p, j0, sg, h->flux.p);
            CK(cudaEventRecord(h->ev_free[b], h->stream));
        }
        else
        {
            if (c >= 2) CK(cudaStreamWaitEvent(h->stream, h->ev_free[b], 0));
            repack_blocks_kernel<false><<<grid, 256, 0, h->stream>>>(cnt, h->ne, h->n, h->np, h->slot_of_block.p, j0, sg, h->flux.p);
            CK(cudaEventRecord(h->ev_ready[b], h->stream));
            CK(cudaStreamWaitEvent(h->xfer, h->ev_ready[b], 0));
            CK(cudaMemcpyAsync(hp, sg, bytes, cudaMemcpyDeviceToHost, h->xfer));
            CK(cudaEventRecord(h->ev_free[b], h->xfer));
        }
        ++h->launches;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaStreamSynchronize(h->xfer));
    return PBTE_OK;
}

int check_handle(pbte_handle *h)
{
    if (!h) return fail(PBTE_ERR_INVALID, "null handle");
    return PBTE_OK;
}

// Launch description of a tail kernel (csrc/tail.cuh).  nparts / per: see DevTail; finish: go on to the end of the iteration;
// the sums go to mom_part when a communicator is attached and the finish does not follow at once (NCCL reduces mom_part -> mom).
DevTail make_tail(pbte_handle *h, int nparts, int per, bool finish)
{
    DevTail Z{};
    Z.nparts = nparts;
    Z.slots_per_chunk = per;
    Z.finish = finish ? 1 : 0;
    Z.m_begin = 0;
    Z.m_end = 1 + h->dim;
    Z.part = h->partial.p;
    Z.mom = (!finish && h->comm) ? h->mom_part.p : h->mom.p;
    Z.tickets = h->tickets.p;
    Z.basis = h->basis.p;
    Z.cell = h->cell.p;
    Z.tv_prev = h->tv_prev.p;
    Z.tc_prev = h->tc_prev.p;
    Z.block_part = h->block_part.p;
    Z.ctrl = h->ctrl.p;
    Z.hist = h->d_hist;
    Z.peer = 0;
    Z.nranks = h->comm_size;
    Z.rank = h->comm_rank;
    Z.qpart = h->mom_part.p;
    Z.xchg = h->xchg.p;
    Z.peers = h->d_peers.p;
    Z.epoch = h->xchg_epoch.p;
    return Z;
}

// The finish step on its own (after the caller's or NCCL's all-reduce of the moment vector).  tc_only: a multi-GPU iteration
// all-reduces and finishes the temperature moment alone (all the next sweep and the stop test need); the heat-flux moments of the
// last iteration follow once, at the next synchronisation point (complete_moments).
int enqueue_finish(pbte_handle *h, bool tc_only = false, bool peer = false)
{
    const DevProblem P = make_dev_problem(h);
    DevTail Z = make_tail(h, 0, 0, true);
    if (tc_only) Z.m_end = 1;
    if (peer) Z.peer = 2;
    iteration_tail_kernel<false><<<h->tail_blocks, kTailThreads, 0, h->stream>>>(P, Z);
    ++h->launches;
    CK(cudaGetLastError());
    if (tc_only) h->qc_pending = true;
    return PBTE_OK;
}

// Heat-flux moments of the last executed iteration: their partial sums still sit in mom_part (iterations after a stop are no-ops).
// reduce = false: the caller has issued the all-reduce already (several handles of one process: one NCCL group).
int complete_moments(pbte_handle *h, bool reduce = true)
{
    if (!h->comm || !h->qc_pending) return PBTE_OK;
    const size_t block = (size_t)h->ne * h->n;
    if (reduce)
    {
        const ncclResult_t r = ncclAllReduce(h->mom_part.p + block, h->mom.p + block, (size_t)h->dim * block, ncclDouble, ncclSum, h->comm, h->stream);
        if (r != ncclSuccess) return fail(PBTE_ERR_CUDA, std::string("ncclAllReduce: ") + ncclGetErrorString(r));
    }
    const DevProblem P = make_dev_problem(h);
    DevTail Z = make_tail(h, 0, 0, true);
    Z.finish = 2;
    Z.m_begin = 1;
    iteration_tail_kernel<false><<<h->tail_blocks, kTailThreads, 0, h->stream>>>(P, Z);
    ++h->launches;
    CK(cudaGetLastError());
    h->qc_pending = false;
    return PBTE_OK;
}

// Enqueue the kernels of one sweep on the handle's stream.  piv: 0 search, 1 search+record, 2 replay.  fuse: the tail kernel goes
// on to the finish step (single-GPU iteration: the whole iteration tail is one launch).
int enqueue_sweep(pbte_handle *h, int piv, bool fuse, int64_t *launch_count, bool peer = false)
{
    const DevProblem P = make_dev_problem(h);
    int64_t launches = 0;
    h->sweep_finished = false;
    if (h->nslots() == 0 || h->max_levels == 0)
    {
        // a rank without directions contributes zeros
        double *dst = (!fuse && h->comm) ? h->mom_part.p : h->mom.p;
        CK(cudaMemsetAsync(dst, 0, h->mom.count * sizeof(double), h->stream));
        if (peer)
        {
            // ... and publishes them like every other rank (the zeroed partial vector is "summed" into the exchange buffer)
            DevTail Z = make_tail(h, 1, 0, false);
            Z.part = h->mom_part.p;
            Z.peer = 1;
            iteration_tail_kernel<false><<<h->tail_blocks, kTailThreads, 0, h->stream>>>(P, Z);
            ++launches;
            CK(cudaGetLastError());
        }
        if (fuse)
        {
            iteration_tail_kernel<false><<<h->tail_blocks, kTailThreads, 0, h->stream>>>(P, make_tail(h, 0, 0, true));
            ++launches;
            CK(cudaGetLastError());
        }
        if (launch_count) *launch_count = launches;
        return PBTE_OK;
    }
    if (h->kernel_used == 3)
    {
        // one marching CTA per slot, then one streaming pass for the moments
        const DevMarch S = make_dev_march(h);
#ifdef PBTE_TRACE
        if (const char *path = std::getenv("PBTE_TRACE_FILE"))
        {
            // development build: per-tile clock stamps of CTA 0 of this sweep, dumped as text (one row per tile)
            const size_t rows_t = (size_t)h->march_ntiles, cnt = rows_t * 16;
            if (!h->m_trace.p) CK(h->m_trace.alloc(cnt));
            CK(cudaMemsetAsync(h->m_trace.p, 0, cnt * sizeof(long long), h->stream));
            DevMarch St = make_dev_march(h);
            if (h->march_regop) pick_regop_kernel(h->n)<<<h->nslots(), regop_threads(h->n), h->march_smem, h->stream>>>(P, St);
            else if (h->march_persist) pick_march_kernel(h->n, h->march_stage)<<<h->march_ncta, march_threads(h->n), h->march_smem, h->stream>>>(P, St);
            else pick_classic_kernel(h->n, h->march_stage)<<<h->nslots(), march_threads(h->n), h->march_smem, h->stream>>>(P, St);
            std::vector<long long> host(cnt);
            CK(cudaMemcpyAsync(host.data(), h->m_trace.p, cnt * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            if (FILE *f = std::fopen(path, "w"))
            {
                long long t0 = 0;
                for (size_t i = 0; i < cnt; ++i)
                    if (host[i] && (!t0 || host[i] < t0)) t0 = host[i];
                for (size_t r = 0; r < rows_t; ++r)
                {
                    bool any = false;
                    for (int k = 0; k < 16; ++k) any = any || host[r * 16 + k];
                    if (!any) continue;
                    std::fprintf(f, "%zu", r);
                    for (int k = 0; k < 11; ++k) std::fprintf(f, " %lld", host[r * 16 + k] ? host[r * 16 + k] - t0 : -1);
                    std::fprintf(f, "\n");
                }
                std::fclose(f);
            }
        }
        else
#endif
        if (h->march_regop) pick_regop_kernel(h->n)<<<h->nslots(), regop_threads(h->n), h->march_smem, h->stream>>>(P, S);
        else if (h->march_persist) pick_march_kernel(h->n, h->march_stage)<<<h->march_ncta, march_threads(h->n), h->march_smem, h->stream>>>(P, S);
        else pick_classic_kernel(h->n, h->march_stage)<<<h->nslots(), march_threads(h->n), h->march_smem, h->stream>>>(P, S);
        {
            // inside a stream capture the record must become an event-record NODE (external flag), else it is only a dependency
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            CK(cudaStreamIsCapturing(h->stream, &cs));
            CK(cudaEventRecordWithFlags(h->ev_march, h->stream, cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault));
        }
        // [Tc | Qc] by one streaming pass over the flux; its last CTAs finish the iteration when `fuse`
        const int chunks = h->tail_chunks, per = (h->nslots() + chunks - 1) / chunks;
        DevTail Z = make_tail(h, chunks, per, fuse || (peer && h->peer_fused));
        if (peer && h->peer_fused)
        {
            // the whole multi-GPU iteration tail in this launch: temperature moment exchanged and finished, heat-flux partial sums
            // left in mom_part for the read-out (complete_moments)
            Z.peer = 3;
            Z.m_end = 1;
            h->sweep_finished = true;
            h->qc_pending = true;
        }
        else if (peer) Z.peer = 1;
        iteration_tail_kernel<true><<<dim3((unsigned)h->tail_blocks, (unsigned)chunks), kTailThreads, 0, h->stream>>>(P, Z);
        CK(cudaGetLastError());
        if (launch_count) *launch_count = 2;
        return PBTE_OK;
    }
    if (h->kernel_used == 2)
    {
        WarpKernelFn fn = pick_warp_kernel(h->n, h->dim, piv);
        const size_t smem = sweep_warp_smem(h->n, h->dim, h->shape);
        const int total = h->level_offsets[h->max_levels];
        if (h->warp_grid > 0)
        {
            // the whole sweep in one launch: resident CTAs walk the level-major item list, upwind neighbours tracked by flags
            CK(cudaMemsetAsync(h->item_done.p, 0, h->item_done.count * sizeof(int32_t), h->stream));
            fn<<<std::min(h->warp_grid, total), kWarpsPerCta * 32, smem, h->stream>>>(P, h->items.p, 0, total, h->item_done.p, h->shape);
            ++launches;
        }
        else
            for (int L = 0; L < h->max_levels; ++L)
            {
                const int begin = h->level_offsets[L], count = h->level_offsets[L + 1] - begin;
                if (count == 0) continue;
                fn<<<count, kWarpsPerCta * 32, smem, h->stream>>>(P, h->items.p, begin, count, nullptr, h->shape);
                ++launches;
            }
    }
    else
    {
        const size_t smem = sweep_generic_smem(h->n, h->F);
        for (int L = 0; L < h->max_levels; ++L)
        {
            const int begin = h->level_offsets[L], count = h->level_offsets[L + 1] - begin;
            if (count == 0) continue;
            sweep_generic_kernel<<<count, 128, smem, h->stream>>>(P, h->items.p, begin);
            ++launches;
        }
    }
    // class partials -> moment vector (ascending class order), then the finish step when `fuse`
    {
        DevTail Z = make_tail(h, h->nclass, 0, fuse);
        if (peer) Z.peer = 1;
        iteration_tail_kernel<false><<<h->tail_blocks, kTailThreads, 0, h->stream>>>(P, Z);
    }
    ++launches;
    CK(cudaGetLastError());
    if (launch_count) *launch_count = launches;
    return PBTE_OK;
}

int run_finalize(pbte_handle *h)
{
    int seg = 1;
    while (seg < h->n && seg < 32) seg <<= 1;
    const size_t threads = (size_t)(1 + h->dim) * h->ne * seg;
    finalize_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, h->stream>>>(h->ne, h->n, 1 + h->dim, seg, h->mom.p, h->basis.p, h->cell.p);
    ++h->launches;
    CK(cudaGetLastError());
    return PBTE_OK;
}

int run_norms(pbte_handle *h, double *diff2, double *tv2, double *tc2)
{
    norms_kernel<<<1, 1024, 0, h->stream>>>(h->ne, (size_t)h->ne * h->n, h->cell.p, h->tv_prev.p, h->mom.p, h->norms.p);
    ++h->launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->h_norms, h->norms.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->h_flag, h->error_flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *diff2 = h->h_norms[0];
    *tv2 = h->h_norms[1];
    *tc2 = h->h_norms[2];
    return PBTE_OK;
}

// Tile schedule, operator storage and operator construction of the marching path.  *used = false when the automatic
// choice declines (tiles too empty to beat the per-element kernels).
int setup_march(pbte_handle *h, const std::vector<uint8_t> &cin, const std::vector<uint8_t> &cbin, bool *used)
{
    *used = false;
    const pbte_host::SharedPlan &plan = h->plan;
    const int nclass = h->nclass, F = h->F, n = h->n, ns = h->nslots(), noc = plan.noclass;
    const pbte_host::MarchSchedule sched = pbte_host::build_march_schedule(plan, nclass, h->h_level.data(), cin.data(), cbin.data());
    h->march_ntiles = sched.ntiles;
    h->march_fill = sched.fill;
    h->march_mask_mismatch = sched.mask_mismatch;
    if (sched.mask_mismatch > 0 && h->class_tol > 1e-11)
    {
        // A direction (almost) tangential to a face: n.Omega changes sign inside an operator class, whose normals agree to class_tol
        // only; the shared operators then drop or invent an inflow term of relative size <= class_tol for those elements.  At the
        // default tolerance (1e-12) that is below the parity bar and the path runs (the count is reported: pbte_stats_t::
        // march_mask_mismatch); with a looser tolerance the per-element kernels take the mesh as it is.
        if (h->kernel_choice == 3 || h->reps_only)
            return fail(PBTE_ERR_INVALID, "shared-operator kernel declined: " + std::to_string(sched.mask_mismatch) +
                                              " (direction, element) pairs see other inflow faces than their operator class (a direction tangential to a face within class_tol > 1e-11); use kernel 0 / 2 with PBTE_IMAGE=full or a smaller class_tol");
        return PBTE_OK;
    }
    if (h->kernel_choice == 0 && sched.fill < 0.4) return PBTE_OK;

    // operator block layout per slot: for every operator class P, then Q_j of the class's inflow faces (ascending j)
    const size_t FR = (size_t)frag_doubles(n), SP = (size_t)split_doubles(n);
    std::vector<int32_t> op_off((size_t)nclass * noc), op_cnt(nclass), slot_class(ns);
    for (int c = 0; c < nclass; ++c)
    {
        int cnt = 0;
        for (int o = 0; o < noc; ++o)
        {
            op_off[(size_t)c * noc + o] = cnt;
            cnt += 1 + __builtin_popcount(sched.cls_inmask[(size_t)c * noc + o]);
        }
        op_cnt[c] = cnt;
    }
    int max_cnt = 0;
    for (int c = 0; c < nclass; ++c)
    {
        max_cnt = std::max(max_cnt, (int)op_cnt[c]);
        for (int s = h->cls_slot_begin[c]; s < h->cls_slot_begin[c + 1]; ++s) slot_class[s] = c;
    }
    int st;
    if ((st = upload(h->m_slot_class, slot_class.data(), slot_class.size(), h->stream))) return st;
    if ((st = upload(h->m_cls_level_begin, sched.cls_level_begin.data(), sched.cls_level_begin.size(), h->stream))) return st;
    if ((st = upload(h->m_level_tile_begin, sched.level_tile_begin.data(), sched.level_tile_begin.size(), h->stream))) return st;
    if ((st = upload(h->m_tile_oclass, sched.tile_oclass.data(), sched.tile_oclass.size(), h->stream))) return st;
    if ((st = upload(h->m_tile_desc, sched.tile_desc.data(), sched.tile_desc.size(), h->stream))) return st;
    if ((st = upload(h->m_cls_inmask, sched.cls_inmask.data(), sched.cls_inmask.size(), h->stream))) return st;
    if ((st = upload(h->m_cls_op_off, op_off.data(), op_off.size(), h->stream))) return st;
    if ((st = upload(h->m_cls_op_cnt, op_cnt.data(), op_cnt.size(), h->stream))) return st;
    {
        // wall temperatures with rows padded to an even count (16-byte rows for the asynchronous copies)
        const int Fp = (F + 1) & ~1;
        std::vector<double> tb((size_t)h->ne * Fp, 0.0);
        for (int e = 0; e < h->ne; ++e)
            for (int j = 0; j < F; ++j) tb[(size_t)e * Fp + j] = plan.ctbc[(size_t)e * F + j];
        if ((st = upload(h->m_ctbc, tb.data(), tb.size(), h->stream))) return st;
        CK(cudaStreamSynchronize(h->stream));
    }
    const size_t nh = (size_t)ns * noc * F * n;
    CK(h->m_oph.alloc(nh));
    CK(cudaMemsetAsync(h->m_oph.p, 0, nh * sizeof(double), h->stream));
    CK(cudaStreamSynchronize(h->stream));  // the host vectors above go out of scope

    // operators staged in shared memory when a slot's block fits (two CTAs per SM up to 100 KB each, one beyond)
    int max_optin = 0;
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    h->march_max_op_cnt = max_cnt;
    h->march_max_class_tiles = sched.max_class_tiles;
    // shared memory of a marching CTA: operators (staged when they fit) + metadata/barriers + tile ring.  Two CTAs per SM when
    // that leaves room for a ring of >= 6 tiles each, one CTA otherwise; without room for operators they are read through L1.
    // staged vectors per ring slot: u_old, T_old and one per inflow face; the kernels whose source producer forms X = c1 T_old + c2 u_old
    // itself (march_xsrc) stage one vector less.  The register-operator kernel keeps the two source vectors.
    const int nv_regop = 2 + sched.max_inflow, nv_march = (march_xsrc(n) ? 1 : 2) + sched.max_inflow;
    h->march_nv = nv_march;
    const size_t h_all = (size_t)noc * F * n;
    // PBTE_MARCH_NOSTAGE=1 forces the large-mesh code path (operators and wall vectors read through L1 instead of staged)
    const bool allow_stage = !(std::getenv("PBTE_MARCH_NOSTAGE") && std::atoi(std::getenv("PBTE_MARCH_NOSTAGE")) != 0);
    h->march_h_smem = allow_stage && h_all * sizeof(double) <= 16 * 1024;
    const size_t h_doubles = h->march_h_smem ? h_all : 0;
    const size_t slot_bytes = march_slot_bytes(n, sched.rows, nv_march, F), slot_bytes_regop = march_slot_bytes(n, sched.rows, nv_regop, F);
    const size_t sm_total = (size_t)max_optin, sm_half = (228 * 1024) / 2 - 2048;
    // Two kernels walk the tiles (sweep_shared.cuh): the persistent one cuts the slots into equal work lists, which pays when there are
    // more slots than resident CTAs AND its operators are staged (measured, profiles/r02_kernel_probes.txt); otherwise the classic one.
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    bool persist = ns > march_ctas_per_sm(n) * sms;
    if (const char *ev = std::getenv("PBTE_MARCH_PERSIST")) persist = std::atoi(ev) != 0;
    auto stage_bytes_of = [&](bool pers) { return (size_t)max_cnt * (pers ? SP : FR) * sizeof(double); };
    auto meta_of = [&](bool pers, int r) {
        return pers ? march_meta_bytes(noc, r, sched.max_class_tiles, h_doubles, march_wpt(n))
                    : march_classic_meta_bytes(sched.max_levels, noc, r, sched.max_class_tiles, h_doubles);
    };
    auto ring_for = [&](bool pers, size_t budget, bool stage, int *ring) {
        const size_t fixed = (stage ? stage_bytes_of(pers) : 0) + 64;
        // mbarrier parity waits are only valid within one phase of the barrier: a slot is always refilled by the same producer
        // warp (ring size a multiple of the producer warps per role: a warp works in order, so the slot's previous tile is loaded
        // before it waits for the slot again) and the ring holds at least one tile per consumer warp (a consumer must find the
        // slot of its next tile at most one phase behind).  A ring that breaks the first rule hangs on large meshes.
        int step = march_src_warps(n);  // a multiple of both loader-warp counts, and even
        while (step % march_nbr_warps(n) != 0 || step % 2 != 0) step += march_src_warps(n);
        int cap = march_ctas_per_sm(n) == 1 ? 2 * march_consumers(n) + 8 : 16;  // deeper than that returns nothing
        if (const char *ev = std::getenv("PBTE_RING_CAP")) cap = std::max(std::atoi(ev), march_min_ring(n));
        for (int r = cap / step * step; r >= march_min_ring(n) && r > 0; r -= step)
            if (fixed + meta_of(pers, r) + (size_t)r * slot_bytes <= budget)
            {
                *ring = r;
                return true;
            }
        return false;
    };
    int ring = 0;
    // One operator class with at most kRegMaxIn inflow faces can use the register-operator kernel (sweep_regop.cuh).  Measured
    // (profiles/README.md): it wins where a slot's operators do not fit the half-SM share of shared memory (hex order 3: 131 KB,
    // 0.28 against 0.37 ms on 512 hexes x 256 slots) and loses ~10 % where they do (hex order 2, quad order 4), so that is the
    // automatic rule; PBTE_MARCH_REGOP=0/1 forces a variant.  Ring: as deep as the CTA's share of shared memory allows, at least
    // two tiles per consumer group.
    h->march_regop = false;
    {
        int unused = 0;
        bool want = !(allow_stage && ring_for(false, sm_half, true, &unused));
        if (const char *ev = std::getenv("PBTE_MARCH_REGOP")) want = std::atoi(ev) != 0;
        if (want && noc == 1 && sched.max_inflow <= kRegMaxIn && pick_regop_kernel(n) != nullptr)
        {
            const size_t budget = regop_ctas(n) == 2 ? sm_half : sm_total - 1024;
            for (int r = 16; r >= 2 * regop_groups(n) && r >= 4; r -= 2)
                if (regop_meta_bytes(n, r, h_doubles) + 16 + (size_t)r * slot_bytes_regop <= budget)
                {
                    h->march_regop = true;
                    ring = r;
                    break;
                }
        }
    }
    MarchKernelFn fn = nullptr;
    h->march_persist = false;
    if (h->march_regop)
    {
        h->march_stage = false;
        h->march_smem = regop_meta_bytes(n, ring, h_doubles) + 16 + (size_t)ring * slot_bytes_regop;
        h->march_nv = nv_regop;
        fn = pick_regop_kernel(n);
    }
    else
    {
        auto configure = [&](bool pers) {
            const size_t first = march_ctas_per_sm(n) == 1 ? sm_total - 1024 : sm_half;
            if (allow_stage && ring_for(pers, first, true, &ring)) h->march_stage = true;
            else if (allow_stage && ring_for(pers, sm_total - 1024, true, &ring)) h->march_stage = true;
            else if (ring_for(pers, first, false, &ring) || ring_for(pers, sm_total - 1024, false, &ring)) h->march_stage = false;
            else return false;
            return true;
        };
        if (!configure(persist)) return fail(PBTE_ERR_INVALID, "marching kernel does not fit in shared memory for this ndof / level count");
        if (persist && !h->march_stage && !std::getenv("PBTE_MARCH_PERSIST"))
        {
            // operators from global memory: the classic kernel is the faster one
            persist = false;
            if (!configure(false)) return fail(PBTE_ERR_INVALID, "marching kernel does not fit in shared memory for this ndof / level count");
        }
        h->march_persist = persist;
        h->march_smem = (h->march_stage ? stage_bytes_of(persist) : 0) + 64 + meta_of(persist, ring) + (size_t)ring * slot_bytes;
        fn = persist ? pick_march_kernel(n, h->march_stage) : pick_classic_kernel(n, h->march_stage);
    }
    h->march_ring = ring;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->march_smem));
    {
        // operator storage: split order where a persistent CTA stages its slot's block, fragment pairs where a classic CTA stages it or
        // the consumers keep it in registers, compact row-major where every tile reads its operators from global memory
        h->march_op_layout = h->march_regop ? 0 : (h->march_stage ? (h->march_persist ? 2 : 0) : (PBTE_COMPACT_OPS == 0 && !h->march_persist ? 0 : 1));
        h->march_op_stride = h->march_op_layout == 0 ? (int)FR : (h->march_op_layout == 2 ? (int)SP : compact_doubles(n));
        std::vector<unsigned long long> slot_base(ns);
        size_t total = 0;
        for (int s = 0; s < ns; ++s)
        {
            slot_base[s] = total;
            total += (size_t)op_cnt[slot_class[s]] * h->march_op_stride;
        }
        if ((st = upload(h->m_slot_op_base, slot_base.data(), slot_base.size(), h->stream))) return st;
        CK(h->m_ops.alloc(total));
        CK(cudaStreamSynchronize(h->stream));
    }

    const size_t build_smem = 2 * (size_t)n * n * sizeof(double);
    if (build_smem > (size_t)max_optin || n > 128) return fail(PBTE_ERR_INVALID, "ndof too large for the operator construction kernel");
    CK(cudaFuncSetAttribute(build_operators_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_smem));
    const DevProblem P = make_dev_problem(h);
    const DevMarch S = make_dev_march(h);
    build_operators_kernel<<<dim3((unsigned)ns, (unsigned)noc), 128, build_smem, h->stream>>>(P, S);
    ++h->launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->h_flag, h->error_flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (*h->h_flag != 0)
    {
        const int flag = *h->h_flag;
        CK(cudaMemsetAsync(h->error_flag.p, 0, sizeof(int32_t), h->stream));
        return fail(flag, "singular local system while building the shared operators");
    }
    // ---- work lists of the persistent marching CTAs (sweep_march_kernel) -------------------------------------------------------
    // The tile sequences of all slots, laid end to end, are cut into equal shares, one per resident CTA; a CTA works through
    // [head of its last slot] [whole slots] [tail of its first slot], so a tail always continues a head that another CTA started with.
    if (h->march_persist)
    {
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, march_threads(n), h->march_smem));
        int resident = std::max(1, per_sm) * std::max(1, sms);
        if (const char *ev = std::getenv("PBTE_MARCH_CTAS")) resident = std::max(1, std::atoi(ev));
        // Cutting slots only pays when there are more slots than resident CTAs: the tiles of a slot depend on each other, so a slot
        // takes the same time whoever processes its parts, and with every slot resident from the start (C2: 256 slots, 296 CTAs) the
        // sweep already ends with the slowest slot (measured: 0.402 ms cut against 0.399 ms whole).  With more slots than CTAs the cut
        // removes the partial last wave (C5 slice: 512 slots, 2.29 -> 2.04 ms; C3 slice 4.77 -> 4.28 ms).
        bool split = (long long)ns > resident;
        if (const char *ev = std::getenv("PBTE_MARCH_SPLIT")) split = std::atoi(ev) != 0;
        std::vector<int32_t> slot_tiles(ns), tile_first(ns);
        long long total = 0;
        for (int sl = 0; sl < ns; ++sl)
        {
            const int c = slot_class[sl], lb = sched.cls_level_begin[c], le = sched.cls_level_begin[c + 1] - 1;
            slot_tiles[sl] = sched.level_tile_begin[le] - sched.level_tile_begin[lb];
            total += slot_tiles[sl];
        }
        const int ncta = split ? (int)std::min<long long>(resident, std::max<long long>(total, 1)) : ns;
        std::vector<int32_t> cta_begin(ncta + 1, 0), pieces;
        if (!split)
            for (int sl = 0; sl < ns; ++sl)
            {
                cta_begin[sl + 1] = sl + 1;
                pieces.insert(pieces.end(), {sl, 0, slot_tiles[sl], slot_tiles[sl]});
            }
        else
        {
            std::vector<long long> off(ns + 1, 0);
            for (int sl = 0; sl < ns; ++sl) off[sl + 1] = off[sl] + slot_tiles[sl];
            int s_lo = 0;
            for (int j = 0; j < ncta; ++j)
            {
                const long long a = total * j / ncta, b = total * (j + 1) / ncta;
                // slots intersecting [a, b)
                while (s_lo < ns && off[s_lo + 1] <= a) ++s_lo;
                std::vector<std::array<int32_t, 4>> mine;
                for (int sl = s_lo; sl < ns && off[sl] < b; ++sl)
                {
                    const int t0 = (int)(std::max(a, off[sl]) - off[sl]), t1 = (int)(std::min(b, off[sl + 1]) - off[sl]);
                    if (t1 > t0) mine.push_back({sl, t0, t1, slot_tiles[sl]});
                }
                // order: head of the last slot first (nothing to wait for), whole slots, the tail of the first slot last
                std::stable_sort(mine.begin(), mine.end(), [](const std::array<int32_t, 4> &x, const std::array<int32_t, 4> &y) {
                    auto rank = [](const std::array<int32_t, 4> &p) { return p[1] > 0 ? 2 : (p[2] < p[3] ? 0 : 1); };
                    return rank(x) < rank(y);
                });
                for (const auto &pc : mine) pieces.insert(pieces.end(), pc.begin(), pc.end());
                cta_begin[j + 1] = (int32_t)(pieces.size() / 4);
            }
        }
        h->march_ncta = ncta;
        h->march_npieces = (int)(pieces.size() / 4);
        if ((st = upload(h->m_cta_piece_begin, cta_begin.data(), cta_begin.size(), h->stream))) return st;
        if ((st = upload(h->m_pieces, pieces.data(), pieces.size(), h->stream))) return st;
        CK(h->m_progress.alloc((size_t)std::max(ns, 1)));
        CK(cudaMemsetAsync(h->m_progress.p, 0, h->m_progress.count * sizeof(int32_t), h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    // tail pass: as few slot chunks as give every SM a streaming CTA.  Measured on C2 (228 element blocks; profiles/r02_kernel_probes.txt,
    // runs j and k): 1 chunk 49 us, 2: 57, 4: 57, 8: 63, 16: 80, 32: 127 us: every extra chunk writes and re-reads a partial moment vector
    {
        int chunks = (sms + h->tail_blocks - 1) / std::max(1, h->tail_blocks);
        chunks = std::max(1, std::min(chunks, std::min(64, std::max(1, ns / 8))));
        if (const char *ev = std::getenv("PBTE_TAIL_CHUNKS")) chunks = std::max(1, std::min(std::atoi(ev), std::max(1, ns)));
        h->tail_chunks = chunks;
    }
    *used = true;
    return PBTE_OK;
}

int alloc_moment_state(pbte_handle *h)
{
    const size_t block = (size_t)h->ne * h->n, nm = 1 + h->dim;
    if (h->np > 2 * kTailThreads) return fail(PBTE_ERR_INVALID, "ndof too large for the iteration tail kernel");
    CK(h->mom.alloc(nm * block));
    CK(h->cell.alloc(nm * h->ne));
    CK(h->tv_prev.alloc(h->ne));
    CK(h->tc_prev.alloc((size_t)h->ne * h->np));
    CK(h->norms.alloc(4));
    CK(h->error_flag.alloc(1));
    CK(cudaMemsetAsync(h->mom.p, 0, nm * block * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->cell.p, 0, nm * h->ne * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->tv_prev.p, 0, h->ne * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->tc_prev.p, 0, h->tc_prev.count * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->error_flag.p, 0, sizeof(int32_t), h->stream));
    // tail kernels: a CTA owns 256 / (np/2) whole elements
    const int epb = kTailThreads / (h->np / 2);
    h->tail_blocks = (h->ne + epb - 1) / epb;
    CK(h->tickets.alloc((size_t)h->tail_blocks + 1));
    CK(cudaMemsetAsync(h->tickets.p, 0, h->tickets.count * sizeof(unsigned), h->stream));
    CK(h->block_part.alloc((size_t)h->tail_blocks * 3));
    CK(h->ctrl.alloc(1));
    CK(cudaMemsetAsync(h->ctrl.p, 0, sizeof(DevCtrl), h->stream));
    CK(cudaHostAlloc(&h->h_hist, sizeof(double) * 4 * kHistRing, cudaHostAllocMapped));
    std::memset(h->h_hist, 0, sizeof(double) * 4 * kHistRing);
    CK(cudaHostGetDevicePointer(&h->d_hist, h->h_hist, 0));
    for (cudaEvent_t &e : h->ev_iter) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return PBTE_OK;
}

// device-side iteration control: tolerance of the stop test, stop flag and iteration counter (all reset together)
int reset_ctrl(pbte_handle *h, double tol)
{
    if (!h->ctrl.p) return PBTE_OK;
    const DevCtrl c{tol, 0, 0};
    CK(cudaMemcpyAsync(h->ctrl.p, &c, sizeof(c), cudaMemcpyHostToDevice, h->stream));  // pageable source: staged before the call returns
    h->iter_enq = 0;
    return PBTE_OK;
}

// One whole source iteration, enqueued: sweep, all-reduce of [Tc | Qc] when a communicator is attached, finish.  No host
// synchronisation; the iteration's history entry becomes valid when the stream reaches the end of the finish step.
int graph_sweep(pbte_handle *h, int mode);
int enqueue_iteration(pbte_handle *h)
{
    int st;
    if (!h->comm) return graph_sweep(h, 1);
    // one process per GPU: the rank's whole iteration (sweep, all-reduce, finish) replays as ONE graph; NCCL calls are capturable
    if (h->peer_mode || (h->use_graph && h->nccl_in_graph)) return graph_sweep(h, 2);
    if ((st = graph_sweep(h, 0))) return st;
    const ncclResult_t r = ncclAllReduce(h->mom_part.p, h->mom.p, (size_t)h->ne * h->n, ncclDouble, ncclSum, h->comm, h->stream);
    if (r != ncclSuccess) return fail(PBTE_ERR_CUDA, std::string("ncclAllReduce: ") + ncclGetErrorString(r));
    if ((st = enqueue_finish(h, true))) return st;
    CK(cudaEventRecord(h->ev2, h->stream));
    ++h->iter_enq;
    return PBTE_OK;
}

// Waits for history entry `index` (an event recorded behind its finish step) and decodes it.
int read_history(pbte_handle *h, int64_t index, cudaEvent_t ev, double *residual)
{
    if (ev) CK(cudaEventSynchronize(ev));
    else CK(cudaStreamSynchronize(h->stream));
    const volatile double *e = h->h_hist + (size_t)(index % kHistRing) * 4;
    const double d2 = e[0], t2 = e[1], c2 = e[2];
    const int flag = (int)e[3];
    h->last_tv_norm = std::sqrt(t2);
    h->last_tc_norm = std::sqrt(c2);
    if (flag != 0)
    {
        cudaMemsetAsync(h->error_flag.p, 0, sizeof(int32_t), h->stream);
        return fail(flag, "a sweep kernel reported a singular local system");
    }
    if (!(t2 > 0.0)) return fail(PBTE_ERR_ZERO_NORM, "Tv norm is zero; cannot compute residual.");
    if (residual) *residual = std::sqrt(d2) / std::sqrt(t2);
    return PBTE_OK;
}

void read_timers(pbte_handle *h)
{
    if (!h->sweep_timed) return;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->last_sweep_ms = ms;
    if (h->kernel_used == 3 && cudaEventElapsedTime(&ms, h->ev0, h->ev_march) == cudaSuccess) h->last_march_ms = ms;
    else h->last_march_ms = 0.0;
    if (cudaEventElapsedTime(&ms, h->ev1, h->ev2) == cudaSuccess) h->last_finish_ms = ms;
    (void)cudaGetLastError();
}
}  // namespace

// =================================================================================================
extern "C" {

int pbte_cuda_abi_version(void) { return PBTE_ABI_VERSION; }
const char *pbte_cuda_last_error(void) { return g_err.c_str(); }

int pbte_cuda_device_count(int *count)
{
    if (!count) return fail(PBTE_ERR_INVALID, "count is null");
    int c = 0;
    const cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess)
    {
        *count = 0;
        return fail(PBTE_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = c;
    return PBTE_OK;
}

void pbte_cuda_default_options(pbte_options_t *opt)
{
    if (!opt) return;
    std::memset(opt, 0, sizeof(*opt));
    opt->device = -1;
    opt->use_graph = 1;
    opt->class_tol = 1e-12;
    opt->max_classes = 512;
}

int pbte_cuda_create(const pbte_problem_t *p, const pbte_options_t *opt_in, pbte_handle_t **out)
{
    if (!p || !out) return fail(PBTE_ERR_INVALID, "null argument");
    NvtxRange nvtx("pbte_cuda_create");
    *out = nullptr;
    if (p->abi_version != PBTE_ABI_VERSION) return fail(PBTE_ERR_INVALID, "problem image has a different ABI version");
    if (p->ne < 0 || p->ndof < 0 || p->ndir < 0 || p->nband < 0 || p->dim < 0 || p->dim > 3 || p->nfaces < 0 || p->nfaces > PBTE_MAX_FACES)
        return fail(PBTE_ERR_INVALID, "invalid sizes in problem image");
    pbte_options_t opt;
    if (opt_in) opt = *opt_in;
    else pbte_cuda_default_options(&opt);
    const auto t_create = std::chrono::steady_clock::now();

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(PBTE_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    int dev = opt.device;
    if (dev < 0) CK(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(PBTE_ERR_INVALID, "device ordinal out of range");

    pbte_handle *h = new (std::nothrow) pbte_handle();
    if (!h) return fail(PBTE_ERR_NOMEM, "out of host memory");
    struct Cleanup
    {
        pbte_handle *h;
        ~Cleanup() { delete h; }
    } cleanup{h};
    h->device = dev;
    DeviceGuard guard(dev);
    if (opt.stream) h->stream = (cudaStream_t)opt.stream;
    else
    {
        CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    CK(cudaEventCreate(&h->ev2));
    CK(cudaEventCreate(&h->ev_march));
    CK(cudaMallocHost(&h->h_norms, 4 * sizeof(double)));
    CK(cudaMallocHost(&h->h_flag, sizeof(int32_t)));
    h->cache_policy = opt.cache_policy;
    h->kernel_choice = opt.kernel;
    h->use_graph = opt.use_graph;
    if (const char *ev = std::getenv("PBTE_NCCL_GRAPH")) h->nccl_in_graph = std::atoi(ev) != 0;
    if (opt.class_tol > 0.0) h->class_tol = opt.class_tol;
    if (opt.max_classes > 0) h->max_classes = opt.max_classes;
    if (opt.kernel < 0 || opt.kernel > 3) return fail(PBTE_ERR_INVALID, "unknown kernel choice");
    h->dim = p->dim;
    h->ne = p->ne;
    h->n = p->ndof;
    h->np = (p->ndof + 1) & ~1;
    h->F = p->nfaces;
    h->ndir = p->ndir;
    h->nband = p->nband;
    h->ms = ((p->ndof * p->ndof + 1) / 2) * 2;
    h->omega = p->omega;
    h->heat_cap_v = p->heat_cap_v;
    h->k0 = opt.dir_begin;
    h->k1 = opt.dir_end;
    if (h->k0 == 0 && h->k1 == 0) h->k1 = p->ndir;
    if (h->k0 < 0 || h->k1 > p->ndir || h->k0 > h->k1) return fail(PBTE_ERR_INVALID, "direction range out of bounds");

    const size_t ne = p->ne, F = p->nfaces, n = p->ndof;
    if (p->dir && p->dir_w && p->ndir > 0)
    {
        h->h_dir.assign(p->dir, p->dir + 3 * (size_t)p->ndir);
        h->h_dirw.assign(p->dir_w, p->dir_w + p->ndir);
        h->has_dirs = true;
    }
    if (p->inv_kn && p->vg && p->heat_cap && p->freq_w && p->nband > 0)
    {
        h->h_invkn.assign(p->inv_kn, p->inv_kn + p->nband);
        h->h_vg.assign(p->vg, p->vg + p->nband);
        h->h_hc.assign(p->heat_cap, p->heat_cap + p->nband);
        h->h_fw.assign(p->freq_w, p->freq_w + p->nband);
        h->nu_max = 0.0;
        for (double v : h->h_invkn) h->nu_max = std::max(h->nu_max, v);  // dt_inv_ (src/PBTESolver.cpp:39-46)
        h->has_bands = true;
    }
    if (p->face_nbr && p->face_normal && ne > 0 && F > 0)
    {
        h->h_fnbr.assign(p->face_nbr, p->face_nbr + ne * F);
        h->h_fnormal.assign(p->face_normal, p->face_normal + ne * F * 3);
        for (int32_t v : h->h_fnbr)
            if (v >= (int32_t)ne || v < PBTE_FACE_NONE) return fail(PBTE_ERR_INVALID, "face_nbr entry out of range");
        int st = upload(h->fnbr, p->face_nbr, ne * F, h->stream);
        if (st) return st;
        st = upload(h->fnormal, p->face_normal, ne * F * 3, h->stream);
        if (st) return st;
        h->has_topology = true;
    }
    if (p->mass || p->stiff || p->face_mass || p->face_coupling || p->face_integral || p->face_tbc)
    {
        if (!(p->mass && p->stiff && p->face_mass && p->face_coupling && p->face_integral && p->face_tbc && h->has_topology && n > 0))
            return fail(PBTE_ERR_INVALID, "incomplete matrix set in problem image");
        int st;
        // operator classes for the shared-operator path (host side, from the caller's arrays)
        if (h->kernel_choice == 0 || h->kernel_choice == 3) h->plan = pbte_host::classify_elements(*p, h->class_tol, h->max_classes);
        // the marching path reads the local matrices of the class representatives only.  The per-element image
        // ((2 + dim + 2F) n^2 doubles per element: 0.4 GB for C2, 17 GB for C4) is uploaded when a per-element kernel may run: kernel
        // 1 / 2 requested, no operator classes found, or an image small enough that keeping the automatic choice open costs nothing
        // (PBTE_IMAGE_LIMIT_MB, default 1024).  PBTE_IMAGE=full / reps forces either.  Without it kernel 0 means kernel 3.
        const double full_mb = (double)(ne * (2 + p->dim + 2 * F) * n * n + ne * F * n) * sizeof(double) / 1048576.0;
        double limit_mb = 1024.0;
        if (const char *ev = std::getenv("PBTE_IMAGE_LIMIT_MB")) limit_mb = std::atof(ev);
        bool reps = h->plan.ok && (h->kernel_choice == 3 || (h->kernel_choice == 0 && full_mb > limit_mb));
        if (const char *ev = std::getenv("PBTE_IMAGE")) reps = h->plan.ok && h->kernel_choice != 1 && h->kernel_choice != 2 && std::string(ev) == "reps";
        if (const char *ev = std::getenv("PBTE_IMAGE"))
            if (std::string(ev) == "full") reps = false;
        h->reps_only = reps;
        if (h->plan.ok && (st = upload_representatives(h, p))) return st;
        if (!reps)
        {
            if ((st = upload_matrices(h->mass, p->mass, ne, (int)n, h->ms, h->stream))) return st;
            {
                std::vector<double> mt(ne * n * n);
                for (size_t e = 0; e < ne; ++e)
                    for (size_t i = 0; i < n; ++i)
                        for (size_t j = 0; j < n; ++j) mt[e * n * n + i + j * n] = p->mass[e * n * n + j + i * n];
                if ((st = upload_matrices(h->massT, mt.data(), ne, (int)n, h->ms, h->stream))) return st;
                CK(cudaStreamSynchronize(h->stream));
            }
            if ((st = upload_matrices(h->stiff, p->stiff, ne * p->dim, (int)n, h->ms, h->stream))) return st;
            if ((st = upload_matrices(h->fmass, p->face_mass, ne * F, (int)n, h->ms, h->stream))) return st;
            if ((st = upload_matrices(h->fcoup, p->face_coupling, ne * F, (int)n, h->ms, h->stream))) return st;
            if ((st = upload(h->fint, p->face_integral, ne * F * n, h->stream))) return st;
            h->image_bytes += (int64_t)((h->mass.count + h->massT.count + h->stiff.count + h->fmass.count + h->fcoup.count + h->fint.count) * sizeof(double));
        }
        else if (h->kernel_choice == 0) h->kernel_choice = 3;
        // a boundary face without a wall temperature (NaN) is reported when a sweep is requested
        if ((st = upload(h->ftbc, p->face_tbc, ne * F, h->stream))) return st;
        for (size_t i = 0; i < ne * F; ++i)
            if (p->face_nbr[i] == PBTE_FACE_BOUNDARY && std::isnan(p->face_tbc[i])) h->missing_bc = true;
        h->has_matrices = true;
    }
    if (p->basis_integrals && ne > 0 && n > 0)
    {
        int st = upload(h->basis, p->basis_integrals, ne * n, h->stream);
        if (st) return st;
        h->has_basis = true;
    }
    if (ne > 0 && n > 0)
    {
        int st = alloc_moment_state(h);
        if (st) return st;
    }
    if (h->has_matrices && h->has_bands && h->has_dirs)
    {
        const size_t fl = ne * (size_t)h->nslots() * h->np;
        CK(h->flux.alloc(fl));
        CK(cudaMemsetAsync(h->flux.p, 0, fl * sizeof(double), h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    h->create_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_create).count();
    cleanup.h = nullptr;
    *out = h;
    return PBTE_OK;
}

int pbte_cuda_destroy(pbte_handle_t *h)
{
    if (!h) return PBTE_OK;
    DeviceGuard guard(h->device);
    cudaStreamSynchronize(h->stream);
    delete h;
    return PBTE_OK;
}

// -------------------------------------------------------------------------------------------------
int pbte_cuda_build_schedule(pbte_handle_t *h)
{
    int st = check_handle(h);
    if (st) return st;
    if (!h->has_topology || !h->has_dirs) return fail(PBTE_ERR_STATE, "build_schedule needs face_nbr, face_normal and directions");
    NvtxRange nvtx("pbte_cuda_build_schedule");
    DeviceGuard guard(h->device);
    const int ne = h->ne, F = h->F, nloc = h->nloc();
    h->scheduled = false;
    h->drop_graphs();
    if (nloc == 0 || ne == 0)
    {
        h->nclass = 0;
        h->max_levels = 0;
        h->level_offsets.assign(1, 0);
        h->scheduled = true;
        return PBTE_OK;
    }
    // 1. upwind masks of every owned direction
    DevBuf<double> d_dirs;
    if ((st = upload(d_dirs, h->h_dir.data() + 3 * (size_t)h->k0, 3 * (size_t)nloc, h->stream))) return st;
    const size_t nm = (size_t)nloc * ne;
    DevBuf<uint8_t> d_in, d_out, d_bin;
    CK(d_in.alloc(nm));
    CK(d_out.alloc(nm));
    CK(d_bin.alloc(nm));
    upwind_mask_kernel<<<(unsigned)((nm + 255) / 256), 256, 0, h->stream>>>(ne, F, h->dim, nloc, d_dirs.p, h->fnormal.p, h->fnbr.p, d_in.p,
                                                                           d_out.p, d_bin.p);
    ++h->launches;
    CK(cudaGetLastError());
    std::vector<uint8_t> in(nm), outm(nm), binm(nm);
    CK(cudaMemcpyAsync(in.data(), d_in.p, nm, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(outm.data(), d_out.p, nm, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(binm.data(), d_bin.p, nm, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));

    // 2. directions with identical upwind sets (interior and wall inflow faces) share one DAG -> one schedule class
    std::map<std::string, int> seen;
    h->class_of_dir.assign(nloc, -1);
    std::vector<int> rep;  // representative direction of each class
    for (int k = 0; k < nloc; ++k)
    {
        std::string key(reinterpret_cast<const char *>(in.data() + (size_t)k * ne), ne);
        key.append(reinterpret_cast<const char *>(binm.data() + (size_t)k * ne), ne);
        auto it = seen.find(key);
        if (it == seen.end())
        {
            it = seen.emplace(std::move(key), (int)rep.size()).first;
            rep.push_back(k);
        }
        h->class_of_dir[k] = it->second;
    }
    const int nclass = h->nclass = (int)rep.size();
    std::vector<int> ndir_in_class(nclass, 0);
    for (int k = 0; k < nloc; ++k) ++ndir_in_class[h->class_of_dir[k]];
    h->cls_slot_begin.assign(nclass + 1, 0);
    for (int c = 0; c < nclass; ++c) h->cls_slot_begin[c + 1] = h->cls_slot_begin[c] + ndir_in_class[c] * std::max(h->nband, 1);
    {
        std::vector<int> fill(nclass, 0);
        h->slotdir_of_dir.assign(nloc, 0);
        for (int k = 0; k < nloc; ++k)
        {
            const int c = h->class_of_dir[k];
            h->slotdir_of_dir[k] = h->cls_slot_begin[c] / std::max(h->nband, 1) + fill[c]++;
        }
    }
    std::vector<uint8_t> cin((size_t)nclass * ne), cbin((size_t)nclass * ne), cout((size_t)nclass * ne, 0);
    for (int c = 0; c < nclass; ++c)
    {
        std::memcpy(cin.data() + (size_t)c * ne, in.data() + (size_t)rep[c] * ne, ne);
        std::memcpy(cbin.data() + (size_t)c * ne, binm.data() + (size_t)rep[c] * ne, ne);
    }
    for (int k = 0; k < nloc; ++k)
    {
        uint8_t *dst = cout.data() + (size_t)h->class_of_dir[k] * ne;
        const uint8_t *src = outm.data() + (size_t)k * ne;
        for (int e = 0; e < ne; ++e) dst[e] |= src[e];
    }
    if ((st = upload(h->cls_in, cin.data(), cin.size(), h->stream))) return st;
    if ((st = upload(h->cls_out, cout.data(), cout.size(), h->stream))) return st;

    // 3. longest-path relaxation of (pass, level) for every class
    DevBuf<int32_t> d_pass, d_level, d_changed;
    const size_t nc = (size_t)nclass * ne;
    CK(d_pass.alloc(nc));
    CK(d_level.alloc(nc));
    CK(d_changed.alloc(1));
    CK(cudaMemsetAsync(d_pass.p, 0, nc * sizeof(int32_t), h->stream));
    CK(cudaMemsetAsync(d_level.p, 0, nc * sizeof(int32_t), h->stream));
    int sweeps_done = 0;
    const int batch = 8;  // relaxation sweeps between convergence checks
    for (;;)
    {
        CK(cudaMemsetAsync(d_changed.p, 0, sizeof(int32_t), h->stream));
        for (int b = 0; b < batch; ++b)
        {
            relax_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, h->stream>>>(ne, F, nclass, h->cls_in.p, h->fnbr.p, d_pass.p, d_level.p, d_changed.p);
            ++h->launches;
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h->h_flag, d_changed.p, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        sweeps_done += batch;
        if (*h->h_flag == 0) break;
        // a DAG on ne nodes has no path longer than ne-1 edges: still changing after that means a cycle
        if (sweeps_done > ne + batch) return fail(PBTE_ERR_SWEEP_STALL, "Angular sweep ordering stalled; check mesh connectivity.");
    }
    h->h_pass.resize(nc);
    h->h_level.resize(nc);
    CK(cudaMemcpyAsync(h->h_pass.data(), d_pass.p, nc * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->h_level.data(), d_level.p, nc * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));

    // 4. work items (class, element) grouped by wavefront level
    int max_level = 0;
    for (int32_t l : h->h_level) max_level = std::max(max_level, (int)l);
    h->max_levels = max_level + 1;
    h->level_offsets.assign(h->max_levels + 1, 0);
    for (int32_t l : h->h_level) ++h->level_offsets[l + 1];
    for (int L = 0; L < h->max_levels; ++L) h->level_offsets[L + 1] += h->level_offsets[L];
    {
        std::vector<int> cursor(h->level_offsets.begin(), h->level_offsets.end() - 1);
        std::vector<int2> items(nc);
        for (int c = 0; c < nclass; ++c)
            for (int e = 0; e < ne; ++e) items[cursor[h->h_level[(size_t)c * ne + e]]++] = make_int2(c, e);
        if ((st = upload(h->items, items.data(), items.size(), h->stream))) return st;
        CK(cudaStreamSynchronize(h->stream));
    }
    {
        std::vector<int32_t> csb(h->cls_slot_begin.begin(), h->cls_slot_begin.end());
        if ((st = upload(h->d_cls_slot_begin, csb.data(), csb.size(), h->stream))) return st;
        CK(cudaStreamSynchronize(h->stream));
    }

    // 5. sweep tables (only when the image carries matrices and bands)
    if (h->has_matrices && h->has_bands)
    {
        const int ns = h->nslots();
        std::vector<double> sdir(3 * (size_t)ns), sv(ns), sc1(ns), sc2(ns), scb(ns), smf(ns), sqf(3 * (size_t)ns);
        for (int k = 0; k < nloc; ++k)
        {
            const double *om = &h->h_dir[3 * (size_t)(h->k0 + k)];
            const double wk = h->h_dirw[h->k0 + k];
            for (int b = 0; b < h->nband; ++b)
            {
                const int s = h->slotdir_of_dir[k] * h->nband + b;
                const double nu = h->h_invkn[b], v = h->h_vg[b], C = h->h_hc[b], dw = h->h_fw[b];
                for (int d = 0; d < 3; ++d) sdir[3 * (size_t)s + d] = om[d];
                sv[s] = v;
                sc1[s] = nu * C / h->omega;          // invKn * Cwp / omega_      (src/PBTESolver.cpp:270)
                sc2[s] = h->nu_max - nu;             // dt_inv_ - invKn           (:276)
                scb[s] = C / h->omega;
                const double f = nu * wk * dw / h->heat_cap_v;  // (MacroscopicQuantities.cpp:115-118)
                smf[s] = f;
                for (int d = 0; d < 3; ++d) sqf[3 * (size_t)s + d] = f * v * om[d];  // (:122-126)
            }
        }
        if ((st = upload(h->slot_dir, sdir.data(), sdir.size(), h->stream))) return st;
        if ((st = upload(h->slot_v, sv.data(), sv.size(), h->stream))) return st;
        if ((st = upload(h->slot_c1, sc1.data(), sc1.size(), h->stream))) return st;
        if ((st = upload(h->slot_c2, sc2.data(), sc2.size(), h->stream))) return st;
        if ((st = upload(h->slot_cb, scb.data(), scb.size(), h->stream))) return st;
        if ((st = upload(h->slot_mf, smf.data(), smf.size(), h->stream))) return st;
        if ((st = upload(h->slot_qf, sqf.data(), sqf.size(), h->stream))) return st;
        CK(cudaStreamSynchronize(h->stream));

        // kernel choice and its shared-memory shape
        int nF = 0, nK = 0;
        for (size_t i = 0; i < cin.size(); ++i)
        {
            nF = std::max(nF, __builtin_popcount(cout[i]));
            nK = std::max(nK, __builtin_popcount(cin[i]));
        }
        h->shape = WarpKernelShape{nF, nK};
        h->kernel_used = 1;
        if (h->kernel_choice != 1 && pick_warp_kernel(h->n, h->dim, 0) != nullptr)
        {
            const size_t smem = sweep_warp_smem(h->n, h->dim, h->shape);
            int max_optin = 0;
            CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
            if (smem <= (size_t)max_optin)
            {
                int per_sm = 1 << 30, sms = 0;
                for (int piv = 0; piv < 3; ++piv)
                {
                    CK(cudaFuncSetAttribute(pick_warp_kernel(h->n, h->dim, piv), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    int occ = 0;
                    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pick_warp_kernel(h->n, h->dim, piv), kWarpsPerCta * 32, smem));
                    per_sm = std::min(per_sm, occ);
                }
                CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
                // one launch per sweep with as many CTAs as are resident at once (PBTE_WARP_PERSIST=0: one launch per wavefront level)
                h->warp_grid = per_sm > 0 ? per_sm * sms : 0;
                if (const char *ev = std::getenv("PBTE_WARP_PERSIST"))
                    if (std::atoi(ev) == 0) h->warp_grid = 0;
                h->kernel_used = 2;
            }
        }
        if (h->kernel_choice == 2 && h->kernel_used != 2) return fail(PBTE_ERR_INVALID, "warp-register kernel not available for this ndof/dim");
        if (h->kernel_used == 1)
        {
            const size_t smem = sweep_generic_smem(h->n, h->F);
            CK(cudaFuncSetAttribute(sweep_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        // shared-operator marching path: forced (kernel 3) or chosen when the tiles are reasonably full
        if (h->kernel_choice == 3 && !h->plan.ok)
            return fail(PBTE_ERR_INVALID, "shared-operator kernel not available: " + (h->plan.why.empty() ? std::string("no plan") : h->plan.why));
        if (h->kernel_choice == 3 && pick_march_kernel(h->n, false) == nullptr)
            return fail(PBTE_ERR_INVALID, "shared-operator kernel not instantiated for this ndof");
        if ((h->kernel_choice == 0 || h->kernel_choice == 3) && h->plan.ok && pick_march_kernel(h->n, false) != nullptr)
        {
            bool used = false;
            if ((st = setup_march(h, cin, cbin, &used))) return st;
            if (used) h->kernel_used = 3;
        }
        // partial moment vectors: one per slot chunk of the tail pass (marching path) or one per upwind-DAG class (per-element kernels)
        {
            const size_t vec = (size_t)(1 + h->dim) * ne * h->n;
            if (h->kernel_used == 3)
            {
                if (h->tail_chunks > 1) CK(h->partial.alloc((size_t)h->tail_chunks * vec));
                else h->partial.release();
                h->items.release();  // the level lists of (class, element) items drive the per-element kernels only
            }
            else
            {
                CK(h->partial.alloc((size_t)nclass * vec));
                if (h->kernel_used == 2 && h->warp_grid > 0) CK(h->item_done.alloc((size_t)nclass * ne));
            }
        }
        h->pivots.release();
        h->pivots_recorded = false;
        if (h->kernel_used == 2 && h->cache_policy == 0) CK(h->pivots.alloc((size_t)ne * ns * h->np));
    }
    h->scheduled = true;
    return PBTE_OK;
}

int pbte_cuda_schedule_info(pbte_handle_t *h, int32_t *n_classes, int32_t *max_levels)
{
    int st = check_handle(h);
    if (st) return st;
    if (!h->scheduled) return fail(PBTE_ERR_STATE, "schedule not built");
    if (n_classes) *n_classes = h->nclass;
    if (max_levels) *max_levels = h->max_levels;
    return PBTE_OK;
}

int pbte_cuda_get_order(pbte_handle_t *h, int32_t k, int32_t *order)
{
    int st = check_handle(h);
    if (st) return st;
    if (!h->scheduled) return fail(PBTE_ERR_STATE, "schedule not built");
    if (k < h->k0 || k >= h->k1 || !order) return fail(PBTE_ERR_INVALID, "direction not owned by this handle");
    // reference list order = elements sorted by (pass, id): counting sort of the device-computed pass numbers
    const int ne = h->ne;
    const int32_t *pass = h->h_pass.data() + (size_t)h->class_of_dir[k - h->k0] * ne;
    int maxp = 0;
    for (int e = 0; e < ne; ++e) maxp = std::max(maxp, (int)pass[e]);
    std::vector<int> off(maxp + 2, 0);
    for (int e = 0; e < ne; ++e) ++off[pass[e] + 1];
    for (int p = 0; p <= maxp; ++p) off[p + 1] += off[p];
    for (int e = 0; e < ne; ++e) order[off[pass[e]]++] = e;
    return PBTE_OK;
}

int pbte_cuda_get_levels(pbte_handle_t *h, int32_t k, int32_t *level)
{
    int st = check_handle(h);
    if (st) return st;
    if (!h->scheduled) return fail(PBTE_ERR_STATE, "schedule not built");
    if (k < h->k0 || k >= h->k1 || !level) return fail(PBTE_ERR_INVALID, "direction not owned by this handle");
    std::memcpy(level, h->h_level.data() + (size_t)h->class_of_dir[k - h->k0] * h->ne, sizeof(int32_t) * h->ne);
    return PBTE_OK;
}

// -------------------------------------------------------------------------------------------------
int pbte_cuda_reset_state(pbte_handle_t *h)
{
    int st = check_handle(h);
    if (st) return st;
    DeviceGuard guard(h->device);
    if (h->flux.p) CK(cudaMemsetAsync(h->flux.p, 0, h->flux.count * sizeof(double), h->stream));
    if (h->mom.p)
    {
        CK(cudaMemsetAsync(h->mom.p, 0, h->mom.count * sizeof(double), h->stream));
        CK(cudaMemsetAsync(h->cell.p, 0, h->cell.count * sizeof(double), h->stream));
        CK(cudaMemsetAsync(h->tv_prev.p, 0, h->tv_prev.count * sizeof(double), h->stream));
        CK(cudaMemsetAsync(h->tc_prev.p, 0, h->tc_prev.count * sizeof(double), h->stream));
        CK(cudaMemsetAsync(h->error_flag.p, 0, sizeof(int32_t), h->stream));
    }
    // no wait: the memsets are ordered on the handle's stream like every later call
    return PBTE_OK;
}

static int coeff_slot(pbte_handle_t *h, int32_t k, int32_t band, int *slot)
{
    if (!h->flux.p) return fail(PBTE_ERR_STATE, "handle holds no flux (problem image without matrices/bands)");
    if (!h->scheduled) return fail(PBTE_ERR_STATE, "schedule not built (flux layout follows the schedule classes)");
    if (k < h->k0 || k >= h->k1 || band < 0 || band >= h->nband) return fail(PBTE_ERR_INVALID, "coefficient block index out of range");
    *slot = h->slotdir_of_dir[k - h->k0] * h->nband + band;
    return PBTE_OK;
}

int pbte_cuda_upload_coeff(pbte_handle_t *h, int32_t k, int32_t band, const double *host)
{
    int st = check_handle(h);
    if (st) return st;
    int slot = 0;
    if ((st = coeff_slot(h, k, band, &slot))) return st;
    if (!host) return fail(PBTE_ERR_INVALID, "null host pointer");
    DeviceGuard guard(h->device);
    // a slot of the flux is exactly one reference block coeff[k][b][s] (ndof x ne, column-major)
    const size_t n = h->n, np = h->np;
    CK(cudaMemcpy2DAsync(h->flux.p + (size_t)slot * h->ne * np, np * sizeof(double), host, n * sizeof(double), n * sizeof(double), h->ne,
                         cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PBTE_OK;
}

int pbte_cuda_download_coeff(pbte_handle_t *h, int32_t k, int32_t band, double *host)
{
    int st = check_handle(h);
    if (st) return st;
    int slot = 0;
    if ((st = coeff_slot(h, k, band, &slot))) return st;
    if (!host) return fail(PBTE_ERR_INVALID, "null host pointer");
    DeviceGuard guard(h->device);
    const size_t n = h->n, np = h->np;
    CK(cudaMemcpy2DAsync(host, n * sizeof(double), h->flux.p + (size_t)slot * h->ne * np, np * sizeof(double), n * sizeof(double), h->ne,
                         cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PBTE_OK;
}

int pbte_cuda_upload_coeff_all(pbte_handle_t *h, const double *host)
{
    int st = check_handle(h);
    if (st) return st;
    int slot = 0;
    if (h->nslots() > 0 && (st = coeff_slot(h, h->k0, 0, &slot))) return st;
    if (!host) return fail(PBTE_ERR_INVALID, "null host pointer");
    DeviceGuard guard(h->device);
    return transfer_all_blocks(h, const_cast<double *>(host), true);
}

int pbte_cuda_download_coeff_all(pbte_handle_t *h, double *host)
{
    int st = check_handle(h);
    if (st) return st;
    int slot = 0;
    if (h->nslots() > 0 && (st = coeff_slot(h, h->k0, 0, &slot))) return st;
    if (!host) return fail(PBTE_ERR_INVALID, "null host pointer");
    DeviceGuard guard(h->device);
    return transfer_all_blocks(h, host, false);
}

int pbte_cuda_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return fail(PBTE_ERR_INVALID, "null pointer");
    *ptr = nullptr;
    if (bytes == 0) return PBTE_OK;
    const cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess)
    {
        (void)cudaGetLastError();
        *ptr = nullptr;
        return fail(e == cudaErrorMemoryAllocation ? PBTE_ERR_NOMEM : PBTE_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    }
    return PBTE_OK;
}

int pbte_cuda_host_free(void *ptr)
{
    if (ptr) cudaFreeHost(ptr);
    return PBTE_OK;
}

int pbte_cuda_upload_moments(pbte_handle_t *h, const double *Tc, const double *Tv)
{
    int st = check_handle(h);
    if (st) return st;
    if (!h->mom.p) return fail(PBTE_ERR_STATE, "handle holds no moment state");
    DeviceGuard guard(h->device);
    const size_t block = (size_t)h->ne * h->n;
    if (Tc)
    {
        // one contiguous host-to-device copy; the padded copy for the sweep is made on the device
        CK(cudaMemcpyAsync(h->mom.p, Tc, block * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpy2DAsync(h->tc_prev.p, h->np * sizeof(double), h->mom.p, h->n * sizeof(double), h->n * sizeof(double), h->ne,
                             cudaMemcpyDeviceToDevice, h->stream));
    }
    if (Tv)
    {
        CK(cudaMemcpyAsync(h->cell.p, Tv, h->ne * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->tv_prev.p, h->cell.p, h->ne * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return PBTE_OK;
}

int pbte_cuda_download_moments(pbte_handle_t *h, double *Tc, double *Qc, double *Tv, double *Qv)
{
    int st = check_handle(h);
    if (st) return st;
    if (!h->mom.p) return fail(PBTE_ERR_STATE, "handle holds no moment state");
    if (h->qc_pending && (Qc || Qv))
        return fail(PBTE_ERR_STATE, "heat-flux moments of the enqueued iterations are not reduced yet: call pbte_cuda_iterate_sync (collective) first");
    DeviceGuard guard(h->device);
    const size_t block = (size_t)h->ne * h->n;
    if (Tc && Qc && Qc == Tc + block) CK(cudaMemcpyAsync(Tc, h->mom.p, (1 + h->dim) * block * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    else
    {
        if (Tc) CK(cudaMemcpyAsync(Tc, h->mom.p, block * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (Qc) CK(cudaMemcpyAsync(Qc, h->mom.p + block, h->dim * block * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    if (Tv && Qv && Qv == Tv + h->ne) CK(cudaMemcpyAsync(Tv, h->cell.p, (size_t)(1 + h->dim) * h->ne * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    else
    {
        if (Tv) CK(cudaMemcpyAsync(Tv, h->cell.p, h->ne * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (Qv) CK(cudaMemcpyAsync(Qv, h->cell.p + h->ne, (size_t)h->dim * h->ne * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return PBTE_OK;
}

// -------------------------------------------------------------------------------------------------
}  // extern "C"

namespace
{
int check_sweep_ready(pbte_handle *h)
{
    if (!h->has_matrices || !h->has_bands) return fail(PBTE_ERR_STATE, "sweep needs the full problem image");
    if (!h->scheduled) return fail(PBTE_ERR_STATE, "sweep before build_schedule");
    if (h->missing_bc) return fail(PBTE_ERR_MISSING_BC, "a boundary face has no isothermal wall temperature");
    return PBTE_OK;
}

// One sweep (and its tail) on the handle's stream, replayed as a CUDA graph where possible.
// One sweep (and what follows it) on the handle's stream, replayed as a CUDA graph where possible.
//   mode 0: sweep + partial moment vector; 1: sweep + fused finish (a whole single-GPU iteration); 2: sweep + NCCL all-reduce of the
//   moment vector + finish (a whole iteration of one rank of a multi-GPU run).
int graph_sweep(pbte_handle *h, int mode)
{
    int st;
    int piv = 0;
    if (h->pivots.p) piv = h->pivots_recorded ? 2 : 1;
    CK(cudaEventRecord(h->ev0, h->stream));
    int64_t launches = 0;
    bool graph_ok = h->use_graph && piv != 1 && h->max_levels > 0 && !h->graph_failed[mode];
#ifdef PBTE_TRACE
    if (std::getenv("PBTE_TRACE_FILE")) graph_ok = false;  // the traced sweep synchronises and writes a file
#endif
    auto enqueue_all = [&]() -> int {
        int rc = enqueue_sweep(h, piv, mode == 1, &launches, mode == 2 && h->peer_mode);
        if (rc || mode != 2) return rc;
        if (h->peer_mode)
        {
            // the finish step reads the peers' temperature partial sums over NVLink itself: no collective call between the kernels
            // (marching path: the tail launch has done that already)
            if (h->sweep_finished) return PBTE_OK;
            if ((rc = enqueue_finish(h, true, true))) return rc;
            ++launches;
            return PBTE_OK;
        }
        // the temperature moment alone: the next sweep and the stop test need nothing else (complete_moments brings the rest)
        const ncclResult_t r = ncclAllReduce(h->mom_part.p, h->mom.p, (size_t)h->ne * h->n, ncclDouble, ncclSum, h->comm, h->stream);
        if (r != ncclSuccess) return fail(PBTE_ERR_CUDA, std::string("ncclAllReduce: ") + ncclGetErrorString(r));
        ++launches;
        if ((rc = enqueue_finish(h, true))) return rc;
        ++launches;
        return PBTE_OK;
    };
    const int g = mode;
    if (graph_ok && (!h->graph_exec[g] || h->graph_mode[g] != piv))
    {
        if (h->graph_exec[g])
        {
            cudaGraphExecDestroy(h->graph_exec[g]);
            h->graph_exec[g] = nullptr;
        }
        cudaGraph_t graph = nullptr;
        const int64_t launches_before = h->launches;
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        st = enqueue_all();
        const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        h->launches = launches_before;  // nothing ran during the capture
        cudaError_t ie = cudaSuccess;
        if (st == PBTE_OK && ce == cudaSuccess) ie = cudaGraphInstantiate(&h->graph_exec[g], graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (st != PBTE_OK || ce != cudaSuccess || ie != cudaSuccess)
        {
            // capture or instantiation failed: this handle goes on with plain launches (a graph with an NCCL node: without NCCL in it)
            (void)cudaGetLastError();
            h->graph_exec[g] = nullptr;
            h->graph_failed[g] = true;
            if (mode == 2) h->nccl_in_graph = false;
            else h->use_graph = 0;
            graph_ok = false;
            // NCCL form: sweep graph, all-reduce and finish as separate enqueues; peer form: the same kernels as plain launches below
            if (mode == 2 && !h->peer_mode) return enqueue_iteration(h);
        }
        else
        {
            h->graph_mode[g] = piv;
            h->launches_per_sweep = launches;
            // the capture consumed the event record: re-record on the live stream
            CK(cudaEventRecord(h->ev0, h->stream));
        }
    }
    if (graph_ok)
    {
        CK(cudaGraphLaunch(h->graph_exec[g], h->stream));
        launches = h->launches_per_sweep;
    }
    else
    {
        launches = 0;
        const int64_t launches_before = h->launches;
        if ((st = enqueue_all())) return st;
        h->launches = launches_before;
        h->launches_per_sweep = launches;
    }
    CK(cudaEventRecord(h->ev1, h->stream));
    if (mode != 0) CK(cudaEventRecord(h->ev2, h->stream));
    h->sweep_timed = true;
    h->launches += launches;
    ++h->sweeps;
    if (mode != 0) ++h->iter_enq;
    if (mode == 2) h->qc_pending = true;
    if (piv == 1) h->pivots_recorded = true;
    return PBTE_OK;
}

// Source iteration over a set of handles that share one communicator set (or a single handle): up to `look` iterations are in
// flight; the host reads the residual history behind them and stops enqueueing at the first residual below tol.  The device's
// own stop test (csrc/tail.cuh) turns whatever was enqueued beyond that iteration into no-ops, so the state that comes back is
// the one the reference leaves: that of the first iteration with res < tol (src/PBTESolver.cpp:313-330).
int solve_handles(pbte_handle **hs, int nh, int32_t max_iter, double tol, int32_t *iters_done, double *residual, double *history,
                  pbte_iteration_callback_t cb, void *user)
{
    NvtxRange nvtx("pbte_cuda_solve");
    int st = PBTE_OK;
    for (int i = 0; i < nh; ++i)
    {
        if ((st = check_sweep_ready(hs[i]))) return st;
        if (!hs[i]->has_basis) return fail(PBTE_ERR_STATE, "solve needs basis_integrals in the problem image");
    }
    int look = 4;
    if (const char *ev = std::getenv("PBTE_LOOKAHEAD")) look = std::atoi(ev);
    look = std::max(1, std::min(look, (int)pbte_handle::kLookMax));
    for (int i = 0; i < nh; ++i)
    {
        DeviceGuard guard(hs[i]->device);
        if ((st = reset_ctrl(hs[i], tol))) return st;
    }
    double res = 0.0;
    int enq = 0, seen = 0;
    while (seen < max_iter && st == PBTE_OK)
    {
        while (enq < max_iter && enq - seen < look && st == PBTE_OK)
        {
            // sweeps of all handles, then their all-reduces as one NCCL group, then the finish steps (peer mode: every handle's
            // whole iteration is enqueued at once, the finish kernels read each other's partial sums over NVLink)
            if (nh > 1 && hs[0]->peer_mode)
            {
                for (int i = 0; i < nh && st == PBTE_OK; ++i)
                {
                    DeviceGuard guard(hs[i]->device);
                    st = graph_sweep(hs[i], 2);
                }
            }
            else if (nh > 1)
            {
                for (int i = 0; i < nh && st == PBTE_OK; ++i)
                {
                    DeviceGuard guard(hs[i]->device);
                    st = graph_sweep(hs[i], 0);
                }
                if (st) break;
                ncclGroupStart();
                ncclResult_t r = ncclSuccess;
                for (int i = 0; i < nh && r == ncclSuccess; ++i)
                    r = ncclAllReduce(hs[i]->mom_part.p, hs[i]->mom.p, (size_t)hs[i]->ne * hs[i]->n, ncclDouble, ncclSum, hs[i]->comm, hs[i]->stream);
                const ncclResult_t r2 = ncclGroupEnd();
                if (r != ncclSuccess || r2 != ncclSuccess)
                {
                    st = fail(PBTE_ERR_CUDA, std::string("ncclAllReduce: ") + ncclGetErrorString(r != ncclSuccess ? r : r2));
                    break;
                }
                for (int i = 0; i < nh && st == PBTE_OK; ++i)
                {
                    DeviceGuard guard(hs[i]->device);
                    if ((st = enqueue_finish(hs[i], true))) break;
                    if (cudaEventRecord(hs[i]->ev2, hs[i]->stream) != cudaSuccess) st = fail(PBTE_ERR_CUDA, "cudaEventRecord failed");
                    ++hs[i]->iter_enq;
                }
            }
            else
            {
                DeviceGuard guard(hs[0]->device);
                st = enqueue_iteration(hs[0]);
            }
            if (st) break;
            {
                DeviceGuard guard(hs[0]->device);
                if (cudaEventRecord(hs[0]->ev_iter[enq % look], hs[0]->stream) != cudaSuccess) st = fail(PBTE_ERR_CUDA, "cudaEventRecord failed");
            }
            ++enq;
        }
        if (st) break;
        {
            // every handle computes the same residual from the same all-reduced vector: read it from the first one
            DeviceGuard guard(hs[0]->device);
            st = read_history(hs[0], seen, hs[0]->ev_iter[seen % look], &res);
        }
        if (st) break;
        if (history) history[seen] = res;
        if (cb) cb(user, seen + 1, res, hs[0]->last_tv_norm, hs[0]->last_tc_norm);
        ++seen;
        if (res < tol) break;
    }
    // drain (iterations enqueued beyond the stop are no-ops on the device) and leave the controls as they were
    for (int i = 0; i < nh; ++i)
    {
        DeviceGuard guard(hs[i]->device);
        cudaStreamSynchronize(hs[i]->stream);
        read_timers(hs[i]);
        hs[i]->last_tv_norm = hs[0]->last_tv_norm;
        hs[i]->last_tc_norm = hs[0]->last_tc_norm;
        const int st2 = reset_ctrl(hs[i], 0.0);
        if (st == PBTE_OK) st = st2;
        cudaStreamSynchronize(hs[i]->stream);
    }
    // heat-flux moments of the last executed iteration (multi-GPU: all-reduced once, here)
    if (nh > 1)
    {
        bool any = false;
        for (int i = 0; i < nh; ++i) any = any || hs[i]->qc_pending;
        if (any)
        {
            ncclGroupStart();
            ncclResult_t r = ncclSuccess;
            for (int i = 0; i < nh && r == ncclSuccess; ++i)
            {
                const size_t block = (size_t)hs[i]->ne * hs[i]->n;
                r = ncclAllReduce(hs[i]->mom_part.p + block, hs[i]->mom.p + block, (size_t)hs[i]->dim * block, ncclDouble, ncclSum, hs[i]->comm, hs[i]->stream);
            }
            const ncclResult_t r2 = ncclGroupEnd();
            if ((r != ncclSuccess || r2 != ncclSuccess) && st == PBTE_OK) st = fail(PBTE_ERR_CUDA, std::string("ncclAllReduce: ") + ncclGetErrorString(r != ncclSuccess ? r : r2));
        }
    }
    for (int i = 0; i < nh; ++i)
    {
        DeviceGuard guard(hs[i]->device);
        const int st2 = complete_moments(hs[i], nh == 1);
        if (st == PBTE_OK) st = st2;
        cudaStreamSynchronize(hs[i]->stream);
    }
    if (iters_done) *iters_done = seen;
    if (residual) *residual = res;
    return st;
}
}  // namespace

extern "C" {

int pbte_cuda_sweep(pbte_handle_t *h)
{
    int st = check_handle(h);
    if (st) return st;
    if ((st = check_sweep_ready(h))) return st;
    DeviceGuard guard(h->device);
    return graph_sweep(h, 0);
}

int pbte_cuda_moment_buffer(pbte_handle_t *h, double **device_ptr, size_t *count)
{
    int st = check_handle(h);
    if (st) return st;
    if (!h->mom.p) return fail(PBTE_ERR_STATE, "handle holds no moment state");
    if (h->comm) return fail(PBTE_ERR_STATE, "a communicator is attached: the library all-reduces the moment vector itself");
    if (device_ptr) *device_ptr = h->mom.p;
    if (count) *count = h->mom.count;
    return PBTE_OK;
}

int pbte_cuda_finish_iteration(pbte_handle_t *h, double *residual)
{
    int st = check_handle(h);
    if (st) return st;
    if (!h->mom.p || !h->has_basis) return fail(PBTE_ERR_STATE, "finish_iteration needs basis_integrals in the problem image");
    DeviceGuard guard(h->device);
    if (h->comm)
    {
        const ncclResult_t r = ncclAllReduce(h->mom_part.p, h->mom.p, h->mom.count, ncclDouble, ncclSum, h->comm, h->stream);
        if (r != ncclSuccess) return fail(PBTE_ERR_CUDA, std::string("ncclAllReduce: ") + ncclGetErrorString(r));
    }
    // cell integrals, norms of the stop test and the hand-over Tv_prev <- Tv, Tc_prev <- Tc (src/PBTESolver.cpp:313-329): one launch
    if ((st = enqueue_finish(h))) return st;
    h->qc_pending = false;
    CK(cudaEventRecord(h->ev2, h->stream));
    const int64_t index = h->iter_enq++;
    st = read_history(h, index, nullptr, residual);
    read_timers(h);
    return st;
}

int pbte_cuda_iterate_async(pbte_handle_t *h)
{
    int st = check_handle(h);
    if (st) return st;
    if ((st = check_sweep_ready(h))) return st;
    if (!h->has_basis) return fail(PBTE_ERR_STATE, "iterate needs basis_integrals in the problem image");
    DeviceGuard guard(h->device);
    return enqueue_iteration(h);
}

int pbte_cuda_iterate_sync(pbte_handle_t *h, int64_t *iterations, double *residual)
{
    int st = check_handle(h);
    if (st) return st;
    DeviceGuard guard(h->device);
    if ((st = complete_moments(h))) return st;  // collective when a communicator is attached
    CK(cudaStreamSynchronize(h->stream));
    read_timers(h);
    if (iterations) *iterations = h->iter_enq;
    if (h->iter_enq == 0) return PBTE_OK;
    return read_history(h, h->iter_enq - 1, nullptr, residual);
}

int pbte_cuda_last_norms(pbte_handle_t *h, double *tv_norm, double *tc_fnorm)
{
    int st = check_handle(h);
    if (st) return st;
    if (tv_norm) *tv_norm = h->last_tv_norm;
    if (tc_fnorm) *tc_fnorm = h->last_tc_norm;
    return PBTE_OK;
}

int pbte_cuda_solve(pbte_handle_t *h, int32_t max_iter, double tol, int32_t *iters_done, double *residual, double *history)
{
    int st = check_handle(h);
    if (st) return st;
    return solve_handles(&h, 1, max_iter, tol, iters_done, residual, history, nullptr, nullptr);
}

int pbte_cuda_solve_ex(pbte_handle_t *const *handles, int32_t n_handles, int32_t max_iter, double tol, int32_t *iters_done, double *residual,
                       double *history, pbte_iteration_callback_t on_iteration, void *user)
{
    if (!handles || n_handles < 1) return fail(PBTE_ERR_INVALID, "solve_ex: no handles");
    std::vector<pbte_handle *> hs(handles, handles + n_handles);
    for (pbte_handle *h : hs)
    {
        if (!h) return fail(PBTE_ERR_INVALID, "null handle");
        if (n_handles > 1 && (!h->comm || h->comm_size != n_handles))
            return fail(PBTE_ERR_STATE, "solve_ex over several handles needs pbte_cuda_comm_attach_all on exactly these handles");
    }
    return solve_handles(hs.data(), n_handles, max_iter, tol, iters_done, residual, history, on_iteration, user);
}

// ---- communication -------------------------------------------------------------------------------
int pbte_cuda_comm_get_unique_id(void *id)
{
    if (!id) return fail(PBTE_ERR_INVALID, "null id");
    static_assert(sizeof(ncclUniqueId) == PBTE_COMM_ID_BYTES, "PBTE_COMM_ID_BYTES must match ncclUniqueId");
    ncclUniqueId uid;
    const ncclResult_t r = ncclGetUniqueId(&uid);
    if (r != ncclSuccess) return fail(PBTE_ERR_CUDA, std::string("ncclGetUniqueId: ") + ncclGetErrorString(r));
    std::memcpy(id, &uid, sizeof(uid));
    return PBTE_OK;
}

// Exchange buffers for the temperature partial sums and their mapping into every rank (csrc/tail.cuh).  local: all handles live in
// this process (direct pointers after cudaDeviceEnablePeerAccess); otherwise the buffers are shared as CUDA IPC handles, which the
// ranks exchange with an ncclAllGather.  Any failure leaves peer_mode off: the all-reduce stays an NCCL call.
static int alloc_exchange(pbte_handle *h)
{
    const size_t xrow = (size_t)h->ne * h->np;
    const int epb = kTailThreads / (h->np / 2), blocks = (h->ne + epb - 1) / epb;  // one flag per element block of the tail grid
    if (const char *ev = std::getenv("PBTE_PEER_FUSED")) h->peer_fused = std::atoi(ev) != 0;
    CK(h->xchg.alloc(2 * xrow + kXchgFlag0 + (size_t)blocks + 16));
    CK(cudaMemsetAsync(h->xchg.p, 0, h->xchg.count * sizeof(double), h->stream));
    CK(h->xchg_epoch.alloc(1));
    CK(cudaMemsetAsync(h->xchg_epoch.p, 0, sizeof(long long), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PBTE_OK;
}

static bool peer_allowed()
{
    const char *ev = std::getenv("PBTE_PEER_ALLREDUCE");
    return !(ev && std::atoi(ev) == 0);
}

static int enable_peer_ipc(pbte_handle *h)
{
    h->peer_mode = false;
    if (!peer_allowed() || h->comm_size < 2 || h->comm_size > kTailThreads) return PBTE_OK;
    int st = alloc_exchange(h);
    if (st) return st;
    const int n = h->comm_size;
    cudaIpcMemHandle_t mine;
    int ok = cudaIpcGetMemHandle(&mine, h->xchg.p) == cudaSuccess ? 1 : 0;
    (void)cudaGetLastError();
    // handles of all ranks (an all-gather of 64 bytes + 1 validity byte per rank through NCCL: device buffers)
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
    DevBuf<unsigned char> d_all;
    CK(d_all.alloc(rec * n));
    std::vector<unsigned char> rec_mine(rec, 0), all(rec * n, 0);
    std::memcpy(rec_mine.data(), &mine, sizeof(mine));
    rec_mine[sizeof(mine)] = (unsigned char)ok;
    CK(cudaMemcpyAsync(d_all.p + rec * h->comm_rank, rec_mine.data(), rec, cudaMemcpyHostToDevice, h->stream));
    const ncclResult_t r = ncclAllGather(d_all.p + rec * h->comm_rank, d_all.p, rec, ncclUint8, h->comm, h->stream);
    if (r != ncclSuccess) return fail(PBTE_ERR_CUDA, std::string("ncclAllGather: ") + ncclGetErrorString(r));
    CK(cudaMemcpyAsync(all.data(), d_all.p, rec * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    bool all_ok = true;
    for (int k = 0; k < n; ++k) all_ok = all_ok && all[rec * k + sizeof(mine)] == 1;
    std::vector<const double *> peers(n, nullptr);
    int opened_ok = all_ok ? 1 : 0;
    if (all_ok)
        for (int k = 0; k < n; ++k)
        {
            if (k == h->comm_rank)
            {
                peers[k] = h->xchg.p;
                continue;
            }
            cudaIpcMemHandle_t hk;
            std::memcpy(&hk, all.data() + rec * k, sizeof(hk));
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, hk, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
            {
                (void)cudaGetLastError();
                opened_ok = 0;
                break;
            }
            h->ipc_opened.push_back(ptr);
            peers[k] = static_cast<const double *>(ptr);
        }
    // every rank must have mapped every buffer, or nobody uses them (one more tiny collective: min over the ranks)
    DevBuf<int> d_flag;
    CK(d_flag.alloc(1));
    CK(cudaMemcpyAsync(d_flag.p, &opened_ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    const ncclResult_t r2 = ncclAllReduce(d_flag.p, d_flag.p, 1, ncclInt, ncclMin, h->comm, h->stream);
    if (r2 != ncclSuccess) return fail(PBTE_ERR_CUDA, std::string("ncclAllReduce: ") + ncclGetErrorString(r2));
    int agreed = 0;
    CK(cudaMemcpyAsync(&agreed, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (agreed != 1)
    {
        for (void *q : h->ipc_opened) cudaIpcCloseMemHandle(q);
        h->ipc_opened.clear();
        (void)cudaGetLastError();
        return PBTE_OK;
    }
    if ((st = upload(h->d_peers, peers.data(), peers.size(), h->stream))) return st;
    CK(cudaStreamSynchronize(h->stream));
    h->peer_mode = true;
    return PBTE_OK;
}

static int enable_peer_local(pbte_handle *const *hs, int n)
{
    for (int i = 0; i < n; ++i) hs[i]->peer_mode = false;
    if (!peer_allowed() || n < 2 || n > kTailThreads) return PBTE_OK;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
        {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, hs[i]->device, hs[j]->device) != cudaSuccess || !can)
            {
                (void)cudaGetLastError();
                return PBTE_OK;
            }
        }
    std::vector<const double *> peers(n, nullptr);
    for (int i = 0; i < n; ++i)
    {
        DeviceGuard guard(hs[i]->device);
        for (int j = 0; j < n; ++j)
            if (j != i)
            {
                const cudaError_t e = cudaDeviceEnablePeerAccess(hs[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                {
                    (void)cudaGetLastError();
                    return PBTE_OK;
                }
                (void)cudaGetLastError();
            }
        const int st = alloc_exchange(hs[i]);
        if (st) return st;
        peers[i] = hs[i]->xchg.p;
    }
    for (int i = 0; i < n; ++i)
    {
        DeviceGuard guard(hs[i]->device);
        const int st = upload(hs[i]->d_peers, peers.data(), peers.size(), hs[i]->stream);
        if (st) return st;
        CK(cudaStreamSynchronize(hs[i]->stream));
    }
    for (int i = 0; i < n; ++i) hs[i]->peer_mode = true;
    return PBTE_OK;
}

static int after_attach(pbte_handle *h)
{
    h->drop_graphs();
    CK(h->mom_part.alloc(h->mom.count));
    CK(cudaMemsetAsync(h->mom_part.p, 0, h->mom.count * sizeof(double), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PBTE_OK;
}

int pbte_cuda_comm_attach(pbte_handle_t *h, const void *id, int32_t rank, int32_t world_size)
{
    int st = check_handle(h);
    if (st) return st;
    if (!id || world_size < 1 || rank < 0 || rank >= world_size) return fail(PBTE_ERR_INVALID, "comm_attach: bad rank / world size");
    if (h->comm) return fail(PBTE_ERR_STATE, "a communicator is already attached");
    if (!h->mom.p) return fail(PBTE_ERR_STATE, "handle holds no moment state");
    DeviceGuard guard(h->device);
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    const ncclResult_t r = ncclCommInitRank(&h->comm, world_size, uid, rank);
    if (r != ncclSuccess)
    {
        h->comm = nullptr;
        return fail(PBTE_ERR_CUDA, std::string("ncclCommInitRank: ") + ncclGetErrorString(r));
    }
    h->comm_rank = rank;
    h->comm_size = world_size;
    if ((st = after_attach(h))) return st;
    return enable_peer_ipc(h);
}

int pbte_cuda_comm_attach_all(pbte_handle_t *const *handles, int32_t n)
{
    if (!handles || n < 1) return fail(PBTE_ERR_INVALID, "comm_attach_all: no handles");
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i)
    {
        if (!handles[i] || !handles[i]->mom.p) return fail(PBTE_ERR_INVALID, "comm_attach_all: null handle or handle without moment state");
        if (handles[i]->comm) return fail(PBTE_ERR_STATE, "a communicator is already attached");
        if (handles[i]->mom.count != handles[0]->mom.count) return fail(PBTE_ERR_INVALID, "comm_attach_all: handles of different problems");
        devs[i] = handles[i]->device;
        for (int j = 0; j < i; ++j)
            if (devs[j] == devs[i]) return fail(PBTE_ERR_INVALID, "comm_attach_all: two handles on one device (NCCL needs one device per rank)");
    }
    std::vector<ncclComm_t> comms(n, nullptr);
    const ncclResult_t r = ncclCommInitAll(comms.data(), n, devs.data());
    if (r != ncclSuccess) return fail(PBTE_ERR_CUDA, std::string("ncclCommInitAll: ") + ncclGetErrorString(r));
    for (int i = 0; i < n; ++i)
    {
        handles[i]->comm = comms[i];
        handles[i]->comm_rank = i;
        handles[i]->comm_size = n;
        DeviceGuard guard(handles[i]->device);
        const int st = after_attach(handles[i]);
        if (st) return st;
    }
    return enable_peer_local(handles, n);
}

int pbte_cuda_comm_detach(pbte_handle_t *h)
{
    int st = check_handle(h);
    if (st) return st;
    DeviceGuard guard(h->device);
    cudaStreamSynchronize(h->stream);
    // graphs that captured the communicator's all-reduce go first: destroying the communicator under them never returns
    h->drop_graphs();
    if (h->comm) ncclCommDestroy(h->comm);
    h->comm = nullptr;
    h->comm_rank = 0;
    h->comm_size = 1;
    for (void *q : h->ipc_opened) cudaIpcCloseMemHandle(q);
    h->ipc_opened.clear();
    h->peer_mode = false;
    h->xchg.release();
    h->mom_part.release();
    h->drop_graphs();
    return PBTE_OK;
}

// -------------------------------------------------------------------------------------------------
int pbte_cuda_macro_reset(pbte_handle_t *h)
{
    int st = check_handle(h);
    if (st) return st;
    if (!h->mom.p) return fail(PBTE_ERR_STATE, "handle holds no moment state");
    DeviceGuard guard(h->device);
    CK(cudaMemsetAsync(h->mom.p, 0, h->mom.count * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->cell.p, 0, h->cell.count * sizeof(double), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PBTE_OK;
}

int pbte_cuda_macro_accumulate(pbte_handle_t *h, int32_t k, int32_t band, const double *coeff)
{
    int st = check_handle(h);
    if (st) return st;
    if (!h->mom.p || !h->has_bands || !h->has_dirs) return fail(PBTE_ERR_STATE, "macro_accumulate needs directions and band tables");
    if (k < 0 || k >= h->ndir || band < 0 || band >= h->nband || !coeff) return fail(PBTE_ERR_INVALID, "macro_accumulate: index out of range");
    DeviceGuard guard(h->device);
    const size_t block = (size_t)h->ne * h->n;
    if (h->scratch.count < block) CK(h->scratch.alloc(block));
    CK(cudaMemcpyAsync(h->scratch.p, coeff, block * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const double f = h->h_invkn[band] * h->h_dirw[k] * h->h_fw[band] / h->heat_cap_v;
    const double *om = &h->h_dir[3 * (size_t)k];
    const double v = h->h_vg[band];
    accumulate_block_kernel<<<(unsigned)((block + 255) / 256), 256, 0, h->stream>>>(block, h->dim, f, f * v * om[0], f * v * om[1], f * v * om[2],
                                                                                   h->scratch.p, h->mom.p);
    ++h->launches;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return PBTE_OK;
}

int pbte_cuda_macro_finalize(pbte_handle_t *h, const double *basis)
{
    int st = check_handle(h);
    if (st) return st;
    if (!h->mom.p) return fail(PBTE_ERR_STATE, "handle holds no moment state");
    DeviceGuard guard(h->device);
    if (basis)
    {
        if ((st = upload(h->basis, basis, (size_t)h->ne * h->n, h->stream))) return st;
        h->has_basis = true;
    }
    if (!h->has_basis) return fail(PBTE_ERR_STATE, "macro_finalize needs basis_integrals");
    if ((st = run_finalize(h))) return st;
    CK(cudaStreamSynchronize(h->stream));
    return PBTE_OK;
}

int pbte_cuda_macro_residual(pbte_handle_t *h, const double *prev_tv, double *residual)
{
    int st = check_handle(h);
    if (st) return st;
    if (!h->mom.p || !prev_tv || !residual) return fail(PBTE_ERR_INVALID, "macro_residual: bad arguments");
    DeviceGuard guard(h->device);
    // cell[0] (= Tv) is compared against the caller's vector; tv_prev is the staging buffer
    DevBuf<double> saved;
    CK(saved.alloc(h->ne));
    CK(cudaMemcpyAsync(saved.p, h->tv_prev.p, h->ne * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(h->tv_prev.p, prev_tv, h->ne * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    double d2 = 0, t2 = 0, c2 = 0;
    st = run_norms(h, &d2, &t2, &c2);
    CK(cudaMemcpyAsync(h->tv_prev.p, saved.p, h->ne * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (st) return st;
    if (!(t2 > 0.0)) return fail(PBTE_ERR_ZERO_NORM, "Tv norm is zero; cannot compute residual.");
    *residual = std::sqrt(d2) / std::sqrt(t2);
    return PBTE_OK;
}

// -------------------------------------------------------------------------------------------------
int pbte_cuda_get_stats(pbte_handle_t *h, pbte_stats_t *s)
{
    int st = check_handle(h);
    if (st) return st;
    if (!s) return fail(PBTE_ERR_INVALID, "null stats");
    std::memset(s, 0, sizeof(*s));
    s->kernel_launches = h->launches;
    s->sweeps = h->sweeps;
    s->last_sweep_ms = h->last_sweep_ms;
    s->last_march_ms = h->last_march_ms;
    s->last_finish_ms = h->last_finish_ms;
    s->flux_bytes = (int64_t)(h->flux.count * sizeof(double));
    s->launches_per_sweep = h->launches_per_sweep;
    s->kernel_used = h->kernel_used;
    s->n_classes = h->nclass;
    s->max_levels = h->max_levels;
    s->n_operator_classes = h->plan.ok ? h->plan.noclass : 0;
    s->march_tiles = h->march_ntiles;
    s->march_fill = h->march_fill;
    s->march_variant = h->kernel_used == 3 && h->march_regop ? 1 : 0;
    s->operator_bytes = (int64_t)((h->m_ops.count + h->m_oph.count) * sizeof(double));
    s->image_bytes = h->image_bytes;
    s->create_ms = h->create_ms;
    s->image_reps_only = h->reps_only ? 1 : 0;
    s->march_mask_mismatch = (int32_t)std::min<int64_t>(h->march_mask_mismatch, INT32_MAX);
    s->march_persistent = h->kernel_used == 3 && h->march_persist ? 1 : 0;
    s->march_ctas = h->kernel_used == 3 ? (h->march_persist ? h->march_ncta : h->nslots()) : 0;
    return PBTE_OK;
}

int pbte_cuda_get_stream(pbte_handle_t *h, void **stream)
{
    int st = check_handle(h);
    if (st) return st;
    if (stream) *stream = (void *)h->stream;
    return PBTE_OK;
}

int pbte_cuda_measure_peaks(int device, double *fp64_fma_tflops, double *fp64_mma_tflops, double *hbm_copy_gbs)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PBTE_ERR_CUDA, "no CUDA device available");
    if (device < 0) CK(cudaGetDevice(&device));
    DeviceGuard guard(device);
    const cudaError_t e = measure_peaks(fp64_fma_tflops, fp64_mma_tflops, hbm_copy_gbs);
    if (e != cudaSuccess) return fail(PBTE_ERR_CUDA, std::string("measure_peaks: ") + cudaGetErrorString(e));
    return PBTE_OK;
}

}  // extern "C"
