// Host-side planning of the shared-operator sweep (operator de-duplication).
//
// 1. classify_elements: finds the distinct local matrices of the problem image (entry-wise comparison, relative
//    tolerance on the max norm) and groups the elements into OPERATOR CLASSES: elements of one class have the same
//    M, G_d and, face by face in a canonical face order (sorted by normal / face-mass id, so the grouping does not
//    depend on the per-element face numbering of ElementIntegralData), the same F_f, n_f, K_f and g_f.  For every
//    (operator class, direction, band) the local system matrix A is therefore the same, and the sweep update
//        u_e = A^-1 [ M^T (c1 T_old + c2 u_old) - sum_in c_f K_f u_nbr(f) - sum_in,bdr c_f (C/omega) T_bc g_f ]
//    becomes u_e = P x_e + sum_f Q_f u_nbr(f) + sum_f T_bc(f) h_f with operators shared by the whole class.
//    Uniformly refined meshes have a handful of classes (1 for Cartesian quads/hexes, tens for refined simplices).
// 2. build_march_schedule: per upwind-DAG class, the wavefront levels cut into tiles of <= 8 elements of one operator
//    class (one DMMA m8n8k4 row block), with the upwind neighbour of every (element, canonical face) resolved.
//
// Nothing here assumes structure; when the mesh has too many classes `ok` stays false and the per-element kernels run.
// Header-only and CUDA-free so that tests/native/plan_check.cpp can exercise it on the CPU.
#pragma once

#include "pbte_cuda.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <queue>
#include <string>
#include <vector>

namespace pbte_host
{
constexpr int kTileElems = 8;     // elements per tile = M of the m8n8k4 MMA
constexpr int kSrcSkip = -1;      // tile source code: face contributes nothing (outflow / padding)
constexpr int kSrcWall = -2;      // tile source code: isothermal wall on an inflow face

// Dictionary of distinct vectors of fixed length under |a - b|_max <= tol * |b|_max.
struct MatDict
{
    size_t len = 0;
    double tol = 1e-12;
    size_t max_reps = 256;
    std::vector<const double *> reps;
    std::vector<double> norm, sum;

    MatDict(size_t len_, double tol_, size_t max_reps_) : len(len_), tol(tol_), max_reps(max_reps_) {}

    // id of the representative equal to m within tol (added when new); -1 when the dictionary is full
    int find_or_add(const double *m)
    {
        double nrm = 0.0, sm = 0.0;
        for (size_t i = 0; i < len; ++i)
        {
            nrm = std::max(nrm, std::fabs(m[i]));
            sm += m[i];
        }
        for (size_t r = 0; r < reps.size(); ++r)
        {
            const double bound = tol * norm[r];
            if (std::fabs(nrm - norm[r]) > bound) continue;
            if (std::fabs(sm - sum[r]) > bound * (double)len) continue;
            const double *q = reps[r];
            bool same = true;
            for (size_t i = 0; i < len; ++i)
                if (std::fabs(m[i] - q[i]) > bound)
                {
                    same = false;
                    break;
                }
            if (same) return (int)r;
        }
        if (reps.size() >= max_reps) return -1;
        reps.push_back(m);
        norm.push_back(nrm);
        sum.push_back(sm);
        return (int)reps.size() - 1;
    }
};

struct SharedPlan
{
    bool ok = false;
    std::string why;  // reason when !ok
    int ne = 0, n = 0, F = 0, dim = 0;
    int noclass = 0;  // operator classes
    int naclass = 0;  // classes of the system matrix alone (diagnostic)
    std::vector<int32_t> elem_oclass;  // [ne]
    std::vector<int32_t> cnbr;         // [ne][F] canonical face j -> neighbour element, PBTE_FACE_BOUNDARY or PBTE_FACE_NONE
    std::vector<uint8_t> cface;        // [ne][F] canonical face j -> local face slot
    std::vector<double> ctbc;          // [ne][F] wall temperature of canonical face j (0 unless boundary)
    std::vector<int32_t> oc_rep;       // [noclass] representative element (source of M, G_d, F_f, n_f)
    std::vector<int32_t> oc_nface;     // [noclass] canonical faces present
    std::vector<int32_t> oc_ksrc;      // [noclass][F] element*F + local face holding K of canonical face j, or -1
    std::vector<int32_t> oc_gsrc;      // [noclass][F] element*F + local face holding g of canonical face j, or -1
};

inline SharedPlan classify_elements(const pbte_problem_t &p, double tol = 1e-12, int max_classes = 512)
{
    SharedPlan out;
    const int ne = out.ne = p.ne, n = out.n = p.ndof, F = out.F = p.nfaces, dim = out.dim = p.dim;
    const size_t nn = (size_t)n * n;
    if (ne == 0 || n == 0 || F == 0)
    {
        out.why = "empty problem";
        return out;
    }
    const size_t max_reps = (size_t)max_classes;
    MatDict dM(nn, tol, max_reps), dG(nn, tol, 4 * max_reps), dF(nn, tol, max_reps), dK(nn, tol, max_reps), dg((size_t)n, tol, max_reps),
        dn(3, tol, 4 * max_reps);
    struct FaceKey
    {
        int nid, fid, f;
        bool operator<(const FaceKey &o) const { return nid != o.nid ? nid < o.nid : (fid != o.fid ? fid < o.fid : f < o.f); }
    };
    std::map<std::vector<int>, int> aclasses;
    std::vector<int32_t> elem_aclass(ne);
    std::vector<int32_t> kid((size_t)ne * F, -1), gid((size_t)ne * F, -1);  // per canonical face
    out.cnbr.assign((size_t)ne * F, PBTE_FACE_NONE);
    out.cface.assign((size_t)ne * F, 0);
    out.ctbc.assign((size_t)ne * F, 0.0);
    std::vector<int> asig;
    std::vector<FaceKey> faces;
    for (int e = 0; e < ne; ++e)
    {
        asig.clear();
        int id = dM.find_or_add(p.mass + e * nn);
        if (id < 0)
        {
            out.why = "too many distinct mass matrices";
            return out;
        }
        asig.push_back(id);
        for (int d = 0; d < dim; ++d)
        {
            id = dG.find_or_add(p.stiff + ((size_t)e * dim + d) * nn);
            if (id < 0)
            {
                out.why = "too many distinct stiffness matrices";
                return out;
            }
            asig.push_back(id);
        }
        faces.clear();
        for (int f = 0; f < F; ++f)
        {
            const size_t slot = (size_t)e * F + f;
            if (p.face_nbr[slot] == PBTE_FACE_NONE) continue;
            const int fid = dF.find_or_add(p.face_mass + slot * nn);
            const int nid = dn.find_or_add(p.face_normal + slot * 3);
            if (fid < 0 || nid < 0)
            {
                out.why = "too many distinct face matrices or normals";
                return out;
            }
            faces.push_back({nid, fid, f});
        }
        std::sort(faces.begin(), faces.end());
        for (size_t j = 0; j < faces.size(); ++j)
        {
            if (j > 0 && faces[j].nid == faces[j - 1].nid && faces[j].fid == faces[j - 1].fid)
            {
                out.why = "two faces of one element share normal and face matrix";
                return out;
            }
            const int f = faces[j].f;
            const size_t slot = (size_t)e * F + f, cs = (size_t)e * F + j;
            asig.push_back(faces[j].nid);
            asig.push_back(faces[j].fid);
            out.cface[cs] = (uint8_t)f;
            out.cnbr[cs] = p.face_nbr[slot];
            if (p.face_nbr[slot] >= 0)
            {
                kid[cs] = dK.find_or_add(p.face_coupling + slot * nn);
                if (kid[cs] < 0)
                {
                    out.why = "too many distinct coupling matrices";
                    return out;
                }
            }
            else
            {
                out.ctbc[cs] = p.face_tbc[slot];
                gid[cs] = dg.find_or_add(p.face_integral + slot * n);
                if (gid[cs] < 0)
                {
                    out.why = "too many distinct face integrals";
                    return out;
                }
            }
        }
        auto it = aclasses.find(asig);
        if (it == aclasses.end())
        {
            if ((int)aclasses.size() >= max_classes)
            {
                out.why = "too many element classes";
                return out;
            }
            it = aclasses.emplace(asig, (int)aclasses.size()).first;
        }
        elem_aclass[e] = it->second;
    }
    const int na = out.naclass = (int)aclasses.size();
    // can all elements of an A-class share one K and one g per canonical face?
    std::vector<int32_t> a_kid((size_t)na * F, -1), a_gid((size_t)na * F, -1);
    std::vector<uint8_t> a_conflict(na, 0);
    for (int e = 0; e < ne; ++e)
    {
        const int a = elem_aclass[e];
        for (int j = 0; j < F; ++j)
        {
            const size_t cs = (size_t)e * F + j, as = (size_t)a * F + j;
            if (kid[cs] >= 0)
            {
                if (a_kid[as] < 0) a_kid[as] = kid[cs];
                else if (a_kid[as] != kid[cs]) a_conflict[a] = 1;
            }
            if (gid[cs] >= 0)
            {
                if (a_gid[as] < 0) a_gid[as] = gid[cs];
                else if (a_gid[as] != gid[cs]) a_conflict[a] = 1;
            }
        }
    }
    // operator classes: the A-class itself, or (A-class, K ids, g ids) where those differ inside the A-class
    std::map<std::vector<int>, int> oclasses;
    out.elem_oclass.resize(ne);
    std::vector<int> osig;
    for (int e = 0; e < ne; ++e)
    {
        const int a = elem_aclass[e];
        osig.assign(1, a);
        if (a_conflict[a])
            for (int j = 0; j < F; ++j)
            {
                osig.push_back(kid[(size_t)e * F + j]);
                osig.push_back(gid[(size_t)e * F + j]);
            }
        auto it = oclasses.find(osig);
        if (it == oclasses.end())
        {
            if ((int)oclasses.size() >= max_classes)
            {
                out.why = "too many operator classes";
                return out;
            }
            it = oclasses.emplace(osig, (int)oclasses.size()).first;
            out.oc_rep.push_back(e);
            int nf = 0;
            for (int j = 0; j < F; ++j)
                if (out.cnbr[(size_t)e * F + j] != PBTE_FACE_NONE) nf = j + 1;
            out.oc_nface.push_back(nf);
            out.oc_ksrc.insert(out.oc_ksrc.end(), F, -1);
            out.oc_gsrc.insert(out.oc_gsrc.end(), F, -1);
        }
        const int o = it->second;
        out.elem_oclass[e] = o;
        for (int j = 0; j < F; ++j)
        {
            const size_t cs = (size_t)e * F + j, os = (size_t)o * F + j;
            const int32_t src = (int32_t)((size_t)e * F + out.cface[cs]);
            if (kid[cs] >= 0 && out.oc_ksrc[os] < 0) out.oc_ksrc[os] = src;
            if (gid[cs] >= 0 && out.oc_gsrc[os] < 0) out.oc_gsrc[os] = src;
        }
    }
    out.noclass = (int)oclasses.size();
    out.ok = true;
    return out;
}

// Per upwind-DAG class: wavefront levels cut into tiles.
struct MarchSchedule
{
    int nclass = 0, F = 0, rows = 0, ntiles = 0, max_levels = 0, max_inflow = 0;  // rows = 2 + F descriptor rows of 8 ints per tile
    int max_class_tiles = 0;               // most tiles of any DAG class
    std::vector<int32_t> cls_level_begin;  // [nclass+1] into level_tile_begin (class c has cls_level_begin[c+1]-cls_level_begin[c]-1 levels)
    std::vector<int32_t> level_tile_begin; // per class: nlevels+1 tile offsets (global tile indices)
    std::vector<int32_t> tile_oclass;      // [ntiles]
    std::vector<int32_t> tile_desc;        // [ntiles][2+F][8]: row 0 = elements (-1 padding), row 1+m = source of the m-th inflow face of the
                                           // tile's class, row 1+F = header {operator class, level, tile index in class, last
                                           // prerequisite tile (index in class; -1 none), 0...}: every upwind neighbour of the tile's
                                           // elements lies in a tile with an index <= that
    std::vector<uint8_t> cls_inmask;       // [nclass][noclass] canonical inflow faces (interior or wall) of the class representative
    double fill = 0.0;                     // elements / (8 * tiles)
    int64_t mask_mismatch = 0;             // (DAG class, element) pairs whose own inflow faces differ from their operator class's: the
                                           // normals of a class agree to the class tolerance only, so n.Omega may change sign inside
                                           // a class for a direction (almost) tangential to a face; the marching path is declined then
};

// Tiles by list scheduling instead of level by level (build_march_schedule, default): a wavefront level of a DAG class rarely holds
// a multiple of 8 elements of each operator class (tetrahedra: 72 classes, ~18 elements per (level, class) at full C3 size and half
// that on a GPU's slice: tiles 52 - 78 % full), and a half-empty tile costs the marching kernel what a full one costs.  An element
// may go into any tile behind the tiles of its upwind neighbours, so tiles are filled across levels: an element becomes ELIGIBLE
// kTileDelay tiles after the last tile of its upwind neighbours (the marching CTA has that many tiles in flight: a tile whose
// prerequisite was emitted just before it would stall the pipeline), and the next tile is taken from the operator class with eight
// eligible elements whose oldest one has waited longest; without a full class, from the class of the oldest eligible element, and
// with nothing eligible from the element that becomes eligible first.  Every DAG class becomes one "level" of the tables.
// Measured on a B200 (profiles/r02_kernel_probes.txt, run h): tetrahedra (C3 slice) 3.78 -> 2.93 ms per sweep at a delay of 8 tiles
// (fill 0.52 -> 0.71; delay 0 fills to 0.80 and takes 7.8 ms: every other tile waits for its predecessor; 4: 3.68, 16: 3.08); Cartesian
// hexes / quads, whose level tiles are 95 % full already, lose 2 - 4 % to the tighter dependencies.  So the level schedule stays where it
// fills its tiles to kMergeBelowFill and list scheduling takes over below (PBTE_TILE_MERGE=0/1 forces one).
constexpr int kTileDelay = 8;
constexpr double kMergeBelowFill = 0.85;

inline void schedule_class_merged(const SharedPlan &plan, const int32_t *lv, const uint8_t *mi, const uint8_t *mb, const uint8_t *inm,
                                  std::vector<int32_t> &tile_desc, std::vector<int32_t> &tile_oclass, int &ntiles, int rows, int delay)
{
    const int ne = plan.ne, F = plan.F, noc = plan.noclass;
    const int t_begin = ntiles;
    // downwind adjacency (reverse edges) and the number of upwind neighbours not yet placed
    std::vector<int32_t> indeg(ne, 0), down_begin(ne + 1, 0), down;
    for (int e = 0; e < ne; ++e)
        for (int j = 0; j < F; ++j)
        {
            const size_t cs = (size_t)e * F + j;
            if (plan.cnbr[cs] >= 0 && (mi[e] >> plan.cface[cs] & 1u))
            {
                ++indeg[e];
                ++down_begin[plan.cnbr[cs] + 1];
            }
        }
    for (int e = 0; e < ne; ++e) down_begin[e + 1] += down_begin[e];
    down.resize(down_begin[ne]);
    {
        std::vector<int32_t> cur(down_begin.begin(), down_begin.end() - 1);
        for (int e = 0; e < ne; ++e)
            for (int j = 0; j < F; ++j)
            {
                const size_t cs = (size_t)e * F + j;
                if (plan.cnbr[cs] >= 0 && (mi[e] >> plan.cface[cs] & 1u)) down[cur[plan.cnbr[cs]]++] = e;
            }
    }
    // per operator class: elements waiting for their eligibility tile (min-heap on that tile), and eligible ones (oldest first)
    typedef std::pair<int64_t, int32_t> Key;  // (tile index or age key, element)
    std::vector<std::priority_queue<Key, std::vector<Key>, std::greater<Key>>> waiting(noc), eligible(noc);
    std::vector<int32_t> tile_of(ne, -1), last_dep(ne, -1);
    auto release = [&](int e) {  // all upwind neighbours of e are placed
        const int64_t at = last_dep[e] < 0 ? 0 : (int64_t)last_dep[e] + 1 + delay;
        waiting[plan.elem_oclass[e]].push({at, e});
    };
    for (int e = 0; e < ne; ++e)
        if (indeg[e] == 0) release(e);
    int placed = 0;
    std::vector<int32_t> members;
    while (placed < ne)
    {
        const int t = ntiles - t_begin;  // index in class of the tile being formed
        int best = -1, best_full = -1;
        int64_t oldest = INT64_MAX, oldest_full = INT64_MAX, soonest = INT64_MAX;
        int soon_class = -1;
        for (int o = 0; o < noc; ++o)
        {
            while (!waiting[o].empty() && waiting[o].top().first <= t)
            {
                const Key k = waiting[o].top();
                waiting[o].pop();
                // age key: wavefront level first (keeps the front compact), then the tile it became eligible at
                eligible[o].push({((int64_t)lv[k.second] << 32) | (int64_t)(uint32_t)k.first, k.second});
            }
            if (!eligible[o].empty())
            {
                const int64_t age = eligible[o].top().first;
                if (age < oldest) oldest = age, best = o;
                if ((int)eligible[o].size() >= kTileElems && age < oldest_full) oldest_full = age, best_full = o;
            }
            if (!waiting[o].empty() && waiting[o].top().first < soonest) soonest = waiting[o].top().first, soon_class = o;
        }
        const int o = best_full >= 0 ? best_full : (best >= 0 ? best : soon_class);
        members.clear();
        while ((int)members.size() < kTileElems && !eligible[o].empty())
        {
            members.push_back(eligible[o].top().second);
            eligible[o].pop();
        }
        if (members.empty())
        {
            // nothing eligible anywhere: the elements of this class that become eligible first go now
            while ((int)members.size() < kTileElems && !waiting[o].empty())
            {
                members.push_back(waiting[o].top().second);
                waiting[o].pop();
            }
        }
        const size_t base = tile_desc.size();
        tile_desc.resize(base + (size_t)rows * kTileElems, kSrcSkip);
        tile_oclass.push_back(o);
        int32_t *hdr = tile_desc.data() + base + (size_t)(1 + F) * kTileElems;
        for (int k = 0; k < kTileElems; ++k) hdr[k] = 0;
        hdr[0] = o;
        hdr[1] = 0;
        hdr[2] = t;
        int dep = -1;
        for (size_t cnt = 0; cnt < members.size(); ++cnt)
        {
            const int e = members[cnt];
            tile_desc[base + cnt] = e;
            int m = 0;
            for (unsigned rest = inm[o]; rest; rest &= rest - 1, ++m)
            {
                const int j = __builtin_ctz(rest);
                const size_t cs = (size_t)e * F + j;
                const int f = plan.cface[cs];
                int32_t code = kSrcSkip;
                if (plan.cnbr[cs] >= 0 && (mi[e] >> f & 1u)) code = plan.cnbr[cs];
                else if (plan.cnbr[cs] == PBTE_FACE_BOUNDARY && (mb[e] >> f & 1u)) code = kSrcWall;
                tile_desc[base + (size_t)(1 + m) * kTileElems + cnt] = code;
                if (code >= 0) dep = std::max(dep, (int)tile_of[code]);
            }
        }
        tile_desc[base + (size_t)(1 + F) * kTileElems + 3] = dep;
        for (int e : members)
        {
            tile_of[e] = t;
            for (int k = down_begin[e]; k < down_begin[e + 1]; ++k)
            {
                const int d = down[k];
                last_dep[d] = std::max(last_dep[d], (int32_t)t);
                if (--indeg[d] == 0) release(d);
            }
        }
        placed += (int)members.size();
        ++ntiles;
    }
}

inline MarchSchedule build_march_schedule_mode(const SharedPlan &plan, int nclass, const int32_t *level, const uint8_t *cin, const uint8_t *cbin, int mode);

// level [nclass][ne], cin / cbin [nclass][ne]: bit f (LOCAL face) = interior / wall inflow face of the class
inline MarchSchedule build_march_schedule(const SharedPlan &plan, int nclass, const int32_t *level, const uint8_t *cin, const uint8_t *cbin)
{
    return build_march_schedule_mode(plan, nclass, level, cin, cbin, -1);
}

// mode: 0 one wavefront level at a time, 1 list scheduling across levels, -1 level schedule unless its tiles are too empty
inline MarchSchedule build_march_schedule_mode(const SharedPlan &plan, int nclass, const int32_t *level, const uint8_t *cin, const uint8_t *cbin, int mode)
{
    if (const char *ev = std::getenv("PBTE_TILE_MERGE")) mode = std::atoi(ev) != 0 ? 1 : 0;
    if (mode < 0)
    {
        MarchSchedule by_level = build_march_schedule_mode(plan, nclass, level, cin, cbin, 0);
        if (by_level.fill >= kMergeBelowFill || std::getenv("PBTE_TILE_MERGE")) return by_level;
        return build_march_schedule_mode(plan, nclass, level, cin, cbin, 1);
    }
    const bool merge = mode == 1;
    int delay = kTileDelay;
    if (const char *ev = std::getenv("PBTE_TILE_DELAY")) delay = std::max(0, std::atoi(ev));
    MarchSchedule S;
    const int ne = plan.ne, F = plan.F;
    S.nclass = nclass;
    S.F = F;
    S.cls_level_begin.assign(nclass + 1, 0);
    S.cls_inmask.assign((size_t)nclass * plan.noclass, 0);
    std::vector<std::pair<int64_t, int32_t>> keyed(ne);
    const int rows = S.rows = 2 + F;
    for (int c = 0; c < nclass; ++c)
    {
        const int32_t *lv = level + (size_t)c * ne;
        const uint8_t *mi = cin + (size_t)c * ne, *mb = cbin + (size_t)c * ne;
        // canonical inflow faces (interior or wall) of every operator class, from its representative's exact masks
        uint8_t *inm = S.cls_inmask.data() + (size_t)c * plan.noclass;
        for (int o = 0; o < plan.noclass; ++o)
        {
            const int e = plan.oc_rep[o];
            unsigned m = 0;
            for (int j = 0; j < F; ++j)
            {
                const size_t cs = (size_t)e * F + j;
                if (plan.cnbr[cs] == PBTE_FACE_NONE) continue;
                const int f = plan.cface[cs];
                if (((mi[e] | mb[e]) >> f) & 1u) m |= 1u << j;
            }
            inm[o] = (uint8_t)m;
            S.max_inflow = std::max(S.max_inflow, (int)__builtin_popcount(m));
        }
        // every element must see exactly the inflow faces its operator class was built for (its operators Q_f, h_f exist for those)
        for (int e = 0; e < ne; ++e)
        {
            unsigned m = 0;
            for (int j = 0; j < F; ++j)
            {
                const size_t cs = (size_t)e * F + j;
                if (plan.cnbr[cs] == PBTE_FACE_NONE) continue;
                if (((mi[e] | mb[e]) >> plan.cface[cs]) & 1u) m |= 1u << j;
            }
            if (m != inm[plan.elem_oclass[e]]) ++S.mask_mismatch;
        }
        if (merge)
        {
            // every element's upwind neighbours must lie in earlier tiles: tiles across levels (schedule_class_merged)
            S.max_levels = std::max(S.max_levels, 1);
            S.cls_level_begin[c + 1] = S.cls_level_begin[c] + 2;
            S.level_tile_begin.resize(S.cls_level_begin[c + 1]);
            const int t_begin = S.ntiles;
            schedule_class_merged(plan, lv, mi, mb, inm, S.tile_desc, S.tile_oclass, S.ntiles, rows, delay);
            S.level_tile_begin[S.cls_level_begin[c]] = t_begin;
            S.level_tile_begin[S.cls_level_begin[c] + 1] = S.ntiles;
            S.max_class_tiles = std::max(S.max_class_tiles, S.ntiles - t_begin);
            continue;
        }
        int nlev = 0;
        for (int e = 0; e < ne; ++e) nlev = std::max(nlev, lv[e] + 1);
        // elements bucketed by level (ascending id inside a level)
        std::vector<int32_t> lev_begin(nlev + 1, 0), by_level(ne);
        for (int e = 0; e < ne; ++e) ++lev_begin[lv[e] + 1];
        for (int L = 0; L < nlev; ++L) lev_begin[L + 1] += lev_begin[L];
        {
            std::vector<int32_t> cur(lev_begin.begin(), lev_begin.end() - 1);
            for (int e = 0; e < ne; ++e) by_level[cur[lv[e]]++] = e;
        }
        S.max_levels = std::max(S.max_levels, nlev);
        S.cls_level_begin[c + 1] = S.cls_level_begin[c] + nlev + 1;
        S.level_tile_begin.resize(S.cls_level_begin[c + 1]);
        int32_t *ltb = S.level_tile_begin.data() + S.cls_level_begin[c];
        const int t_begin = S.ntiles;
        std::vector<int32_t> tile_of(ne, -1);  // tile (index in class) that holds an element, known for the finished levels
        for (int L = 0; L < nlev; ++L)
        {
            ltb[L] = S.ntiles;
            // Order inside a level: operator class, then the LAST tile any upwind neighbour lies in, then id.  The marching kernel
            // starts a tile once every tile up to its last prerequisite is stored, so tiles whose prerequisites finish early come
            // first and the next level overlaps the tail of this one (any order inside a level gives the same numbers: a sweep
            // depends on the DAG only).
            const int cnt_l = lev_begin[L + 1] - lev_begin[L];
            for (int k = 0; k < cnt_l; ++k)
            {
                const int e = by_level[lev_begin[L] + k];
                int dep = -1;
                for (int j = 0; j < F; ++j)
                {
                    const size_t cs = (size_t)e * F + j;
                    if (plan.cnbr[cs] >= 0 && (mi[e] >> plan.cface[cs] & 1u)) dep = std::max(dep, (int)tile_of[plan.cnbr[cs]]);
                }
                keyed[k] = {((int64_t)plan.elem_oclass[e] << 32) | (int64_t)(uint32_t)(dep + 1), e};
            }
            std::sort(keyed.begin(), keyed.begin() + cnt_l);
            for (int i = 0; i < cnt_l;)
            {
                const int o = plan.elem_oclass[keyed[i].second];
                int cnt = 0;
                const size_t base = S.tile_desc.size();
                S.tile_desc.resize(base + (size_t)rows * kTileElems, kSrcSkip);
                S.tile_oclass.push_back(o);
                {
                    int32_t *hdr = S.tile_desc.data() + base + (size_t)(1 + F) * kTileElems;
                    for (int k = 0; k < kTileElems; ++k) hdr[k] = 0;
                    hdr[0] = o;
                    hdr[1] = L;
                    hdr[2] = S.ntiles - t_begin;
                }
                int dep = -1;
                while (i < cnt_l && cnt < kTileElems && plan.elem_oclass[keyed[i].second] == o)
                {
                    const int e = keyed[i].second;
                    S.tile_desc[base + cnt] = e;
                    // row 1+m = m-th inflow face of the operator class (ascending canonical face index): the kernel reads the
                    // rows without knowing the class
                    int m = 0;
                    for (unsigned rest = inm[o]; rest; rest &= rest - 1, ++m)
                    {
                        const int j = __builtin_ctz(rest);
                        const size_t cs = (size_t)e * F + j;
                        const int f = plan.cface[cs];
                        int32_t code = kSrcSkip;
                        if (plan.cnbr[cs] >= 0 && (mi[e] >> f & 1u)) code = plan.cnbr[cs];
                        else if (plan.cnbr[cs] == PBTE_FACE_BOUNDARY && (mb[e] >> f & 1u)) code = kSrcWall;
                        S.tile_desc[base + (size_t)(1 + m) * kTileElems + cnt] = code;
                        if (code >= 0) dep = std::max(dep, (int)tile_of[code]);
                    }
                    ++cnt;
                    ++i;
                }
                // header word 3: last prerequisite tile (index in class; -1 none)
                S.tile_desc[base + (size_t)(1 + F) * kTileElems + 3] = dep;
                ++S.ntiles;
            }
            // the tiles of this level become visible to the next one
            for (int t = ltb[L]; t < S.ntiles; ++t)
                for (int k = 0; k < kTileElems; ++k)
                {
                    const int e = S.tile_desc[(size_t)t * rows * kTileElems + k];
                    if (e >= 0) tile_of[e] = t - t_begin;
                }
        }
        ltb[nlev] = S.ntiles;
        S.max_class_tiles = std::max(S.max_class_tiles, S.ntiles - t_begin);
    }
    S.fill = S.ntiles ? (double)nclass * ne / ((double)kTileElems * S.ntiles) : 0.0;
    return S;
}
}  // namespace pbte_host
