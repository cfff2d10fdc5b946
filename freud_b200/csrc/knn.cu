// k-nearest-neighbour query on the cell list.
//
// Replaces AABBQueryIterator::next (freud/locality/AABBQuery.cc:152-281).  By E3 that
// iterator returns the k smallest closest-image distances in the IMAGE arithmetic r = p_j - (q + image_k),
// restricted to r_min <= d and r_sq < r_max^2, independent of r_guess/scale; ties at the k-th place are
// resolved by an unstable std::sort upstream (unspecified) and by (r_sq, point index) here.
// WRAP flavour = LinkCellQueryIterator::next (freud/locality/LinkCell.cc:575-679): the k smallest wrapped
// distances r = Box::wrap(p_j - q) with r_min^2 <= r_sq < r_max^2; its shell-by-shell early exit only stops once
// the k-th distance is inside the searched shells, so the answer does not depend on the cell width.
//
// One thread per query point keeps its k best candidates in a [k][n_query] scratch array (slot-major, so
// the threads of a warp touch consecutive words).  A query is resolved when its k-th distance is inside the
// radius the visited cells are guaranteed to cover; the host widens the grid and re-runs otherwise.
#include "internal.h"

namespace fgpu {

namespace {

constexpr int kKnnThreads = 128;

struct Cand
{
    float r_sq;
    uint32_t slot;
    uint32_t j;
};

__device__ __forceinline__ bool cand_less(float a_rsq, uint32_t a_j, float b_rsq, uint32_t b_j)
{
    return a_rsq < b_rsq || (a_rsq == b_rsq && a_j < b_j);
}

__global__ void __launch_bounds__(kKnnThreads) k_knn(KnnArgs a)
{
    uint32_t const t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_query)
    {
        return;
    }
    const BoxDev& box = a.box;
    const GridDev& g = a.grid;
    float4 const q = a.q_sorted[t];
    uint32_t const qi = __float_as_uint(q.w);
    uint32_t const q_global = qi + a.q_index_offset;
    float const qx = q.x, qy = q.y, qz = box.is2d ? 0.0f : q.z; // AABBQuery.cc:84-87
    float const r_max_sq = __fmul_rn(a.r_max, a.r_max);
    float const r_min_sq = __fmul_rn(a.r_min, a.r_min); // WRAP flavour: LinkCell.cc:577-578
    float const r_safe_sq = a.r_safe * a.r_safe;
    uint32_t const nq = a.n_query, k = a.k;
    float* const best_d = a.knn_d + qi;
    uint32_t* const best_s = a.knn_s + qi;
    uint32_t count = 0;
    unsigned long long evals = 0;
    bool const any_shift = *g.any_shift_flag != 0;

    int cx, cy, cz, nqx, nqy, nqz;
    cell_coords(box, g.dx, g.dy, g.dz, qx, qy, qz, cx, cy, cz, nqx, nqy, nqz);
    AxisSlots sx, sy, sz;
    make_slots(g.dx, g.amb_x, cx, sx);
    make_slots(g.dy, g.amb_y, cy, sy);
    make_slots(g.dz, g.amb_z, cz, sz);
    for (int iz = 0; iz < sz.n; ++iz)
    {
        for (int iy = 0; iy < sy.n; ++iy)
        {
            for (int ix = 0; ix < sx.n; ++ix)
            {
                uint32_t const cell = ((uint32_t) sz.cell[iz] * g.dy + sy.cell[iy]) * g.dx + sx.cell[ix];
                uint32_t const beg = __ldg(g.cell_start + cell), end = __ldg(g.cell_start + cell + 1);
                int const wx = sx.w[ix], wy = sy.w[iy], wz = sz.w[iz];
                for (uint32_t s = beg; s < end; ++s)
                {
                    float4 const p = __ldg(g.sorted + s);
                    uint32_t const j = __float_as_uint(p.w);
                    if (a.exclude_ii && j == q_global)
                    {
                        continue;
                    }
                    float best = INFINITY;
                    if (a.flavour == FGPU_FLAVOUR_WRAP)
                    {
                        // LinkCellQueryIterator::next: one wrapped displacement per point, LinkCell.cc:617-622
                        ++evals;
                        float rx, ry, rz;
                        wrap_exact(box, __fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y), __fsub_rn(p.z, q.z), rx, ry, rz);
                        best = dot_exact(rx, ry, rz);
                        if (!(best < r_max_sq) || best < r_min_sq)
                        {
                            continue;
                        }
                    }
                    else
                    {
                        int njx = 0, njy = 0, njz = 0;
                        if (any_shift)
                        {
                            unpack_shift(__ldg(g.shift + s), njx, njy, njz);
                        }
                        int const kx0 = wx == 2 ? -1 : njx - nqx - wx, kx1 = wx == 2 ? 1 : kx0;
                        int const ky0 = wy == 2 ? -1 : njy - nqy - wy, ky1 = wy == 2 ? 1 : ky0;
                        int const kz0 = wz == 2 ? -1 : njz - nqz - wz, kz1 = wz == 2 ? 1 : kz0;
                        float const pz = box.is2d ? 0.0f : p.z; // AABBQuery.cc:118-122
                        // closest admissible image (AABBQuery.cc:197-211 keeps the closest image per point)
                        for (int kx = max(kx0, -1); kx <= min(kx1, 1); ++kx)
                        {
                            for (int ky = max(ky0, -1); ky <= min(ky1, 1); ++ky)
                            {
                                for (int kz = max(kz0, -1); kz <= min(kz1, 1); ++kz)
                                {
                                    ++evals;
                                    float ix_, iy_, iz_;
                                    image_vector(box, kx, ky, kz, ix_, iy_, iz_);
                                    float const rx = __fsub_rn(p.x, __fadd_rn(qx, ix_));
                                    float const ry = __fsub_rn(p.y, __fadd_rn(qy, iy_));
                                    float const rz = __fsub_rn(pz, __fadd_rn(qz, iz_));
                                    best = fminf(best, dot_exact(rx, ry, rz));
                                }
                            }
                        }
                        if (!(best < r_max_sq))
                        {
                            continue; // the ball query inside the iterator keeps r_sq < min(r_cur, r_max)^2
                        }
                        if (__fsqrt_rn(best) < a.r_min)
                        {
                            continue; // AABBQuery.cc:213-216
                        }
                    }
                    if (!a.cover_all && best > r_safe_sq)
                    {
                        continue; // cannot be part of a resolved answer
                    }
                    // insert into the sorted top-k
                    uint32_t pos;
                    if (count < k)
                    {
                        pos = count++;
                    }
                    else
                    {
                        float const last = best_d[(size_t) (k - 1) * nq];
                        uint32_t const last_j = __float_as_uint(g.sorted[best_s[(size_t) (k - 1) * nq]].w);
                        if (!cand_less(best, j, last, last_j))
                        {
                            continue;
                        }
                        pos = k - 1;
                    }
                    while (pos > 0)
                    {
                        float const prev = best_d[(size_t) (pos - 1) * nq];
                        uint32_t const prev_s = best_s[(size_t) (pos - 1) * nq];
                        if (prev < best)
                        {
                            break;
                        }
                        if (prev == best && __float_as_uint(g.sorted[prev_s].w) < j)
                        {
                            break;
                        }
                        best_d[(size_t) pos * nq] = prev;
                        best_s[(size_t) pos * nq] = prev_s;
                        --pos;
                    }
                    best_d[(size_t) pos * nq] = best;
                    best_s[(size_t) pos * nq] = s;
                }
            }
        }
    }
    bool const resolved = a.cover_all || a.r_max <= a.r_safe || (count >= k); // kept entries are all <= r_safe
    a.row_counts[qi] = count;
    if (!resolved)
    {
        atomicAdd(a.unresolved, 1ULL);
    }
    atomicAdd(a.total, (unsigned long long) count);
    if (a.evals != nullptr)
    {
        atomicAdd(a.evals, evals);
    }
}

// One thread per (query, kept slot): recompute the closest image in the reference's image order, rank the
// entry inside its row by (j) or (d, j), write the NeighborList arrays.
__global__ void __launch_bounds__(256) k_knn_emit(KnnEmitArgs a)
{
    uint64_t const idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t const qi = (uint32_t) (idx % a.n_query);
    uint32_t const m = (uint32_t) (idx / a.n_query);
    if (m >= a.k)
    {
        return;
    }
    uint32_t const count = a.row_counts[qi];
    if (m >= count)
    {
        return;
    }
    uint32_t const nq = a.n_query;
    uint32_t const s = a.knn_s[(size_t) m * nq + qi];
    float4 const p = __ldg(a.sorted + s);
    uint32_t const j = __float_as_uint(p.w);
    float const qx = a.q_xyz[3 * (size_t) qi], qy = a.q_xyz[3 * (size_t) qi + 1];
    float const qz = a.box.is2d ? 0.0f : a.q_xyz[3 * (size_t) qi + 2];
    float const pz = a.box.is2d ? 0.0f : p.z;
    float rx = 0, ry = 0, rz = 0, r_sq = INFINITY;
    if (a.flavour == FGPU_FLAVOUR_WRAP)
    {
        wrap_exact(a.box, __fsub_rn(p.x, qx), __fsub_rn(p.y, qy), __fsub_rn(p.z, a.q_xyz[3 * (size_t) qi + 2]), rx, ry,
                   rz);
        r_sq = dot_exact(rx, ry, rz);
    }
    for (int code = 0; a.flavour != FGPU_FLAVOUR_WRAP && code < 27; ++code)
    {
        int i = 0, jj = 0, k = 0;
        if (code > 0)
        {
            int const c = code - 1 + (code - 1 >= 13 ? 1 : 0);
            i = c / 9 - 1;
            jj = (c / 3) % 3 - 1;
            k = c % 3 - 1;
        }
        if (a.box.is2d && k != 0)
        {
            continue;
        }
        float ix_, iy_, iz_;
        image_vector(a.box, i, jj, k, ix_, iy_, iz_);
        float const tx = __fsub_rn(p.x, __fadd_rn(qx, ix_));
        float const ty = __fsub_rn(p.y, __fadd_rn(qy, iy_));
        float const tz = __fsub_rn(pz, __fadd_rn(qz, iz_));
        float const t_sq = dot_exact(tx, ty, tz);
        if (t_sq < r_sq)
        {
            rx = tx;
            ry = ty;
            rz = tz;
            r_sq = t_sq;
        }
    }
    float const d = __fsqrt_rn(r_sq);
    // rank inside the row
    uint32_t rank = 0;
    for (uint32_t o = 0; o < count; ++o)
    {
        if (o == m)
        {
            continue;
        }
        uint32_t const oj = __float_as_uint(__ldg(a.sorted + a.knn_s[(size_t) o * nq + qi]).w);
        bool less;
        if (a.sort_by_distance)
        {
            float const od = __fsqrt_rn(a.knn_d[(size_t) o * nq + qi]);
            less = od < d || (od == d && oj < j);
        }
        else
        {
            less = oj < j;
        }
        rank += less ? 1U : 0U;
    }
    uint64_t const out = (uint64_t) a.row_start[qi] + rank;
    a.neighbors[2 * out] = qi;
    a.neighbors[2 * out + 1] = j;
    a.distances[out] = d;
    a.weights[out] = 1.0f;
    a.vectors[3 * out] = rx;
    a.vectors[3 * out + 1] = ry;
    a.vectors[3 * out + 2] = rz;
}

} // namespace

void launch_knn(fgpu_ctx* ctx, const KnnArgs& a)
{
    if (a.n_query == 0)
    {
        return;
    }
    {
        KernelScope ks(ctx, "knn");
        k_knn<<<(a.n_query + kKnnThreads - 1) / kKnnThreads, kKnnThreads, 0, ctx->stream>>>(a);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_knn_emit(fgpu_ctx* ctx, const KnnEmitArgs& a)
{
    uint64_t const total = (uint64_t) a.n_query * a.k;
    if (total == 0)
    {
        return;
    }
    {
        KernelScope ks(ctx, "knn_emit");
        k_knn_emit<<<(unsigned) ((total + 255) / 256), 256, 0, ctx->stream>>>(a);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

} // namespace fgpu
