// BVHAccelerator::construct (reference shapes/BVHAccelerator.h:175-209) on the device: spcu_build_bvh, spcu_triangle_bounds.
//
// The reference builds its tree by single-threaded recursion: fold the primitives' bounds in range order, split the range
// with std::partition at the centre of the node's bounds along its widest axis, recurse left then right.  Primitive IDs are
// positions in the final order, and the flattener numbers internal nodes in depth-first pre-order, so both the ORDER the
// partition leaves behind and the exact bounds (down to the sign of a zero) are part of the contract.  None of it needs the
// recursion:
//
//   * level-synchronous: all nodes of one depth are processed by a handful of passes over the n primitive positions
//     (every position knows the node it currently belongs to), ~10 launches per level, no recursion, no per-node launch;
//     the pass that applies a level's partition also reduces the bounds of the next level's nodes (k_bounds<true>);
//   * the bounds fold is order-dependent only through ties (_mm_min_ps / _mm_max_ps return their SECOND operand on a tie,
//     which shows in +0 / -0): "the last position among the equal minima wins".  A 64-bit key (ordered float | position |
//     sign) under atomicMin / atomicMax reproduces that with no ordered traversal; warps reduce their contiguous runs with
//     shuffles first, and a run that IS the whole node stores without an atomic;
//   * libstdc++'s std::partition (Hoare: swap the first misplaced element from the left with the first from the right) is a
//     closed-form permutation: with k = number of elements satisfying the predicate, the j-th "false" among positions
//     [first, k) changes places with the j-th "true" from the right among [k, last).  One exclusive prefix sum of the
//     predicate over all positions gives every element its partner;
//   * pre-order index of an internal node = (internal nodes whose range starts further left) + (its depth in the chain of
//     left children starting at the same position) — a prefix sum over range starts, no tree walk.
//
// Compiled WITHOUT fast-math and with --fmad=false (Makefile): (lo + hi) / 2 and hi - lo are IEEE operations.
#include "build_util.cuh"

using namespace spcu;

namespace {

constexpr uint32_t kInvalid   = 0xFFFFFFFFu;
constexpr uint32_t kPending   = 0;    // node states
constexpr uint32_t kLeaf      = 1;
constexpr uint32_t kInternal  = 2;
constexpr uint32_t kSplitting = 3;    // more than 4 primitives: partition decides between internal and leaf

// Reduction targets of one node for one level: 3 lower-bound keys (atomicMin), 3 upper-bound keys (atomicMax) and, per axis,
// the first position whose primitive is NOT the degenerate interval [+-0, +-0] (see decode_bounds).
struct NodeKeys
{
    unsigned long long k[6];
    uint32_t           first_nonzero[3];
    uint32_t           pad;
};
static_assert(sizeof(NodeKeys) == 64, "NodeKeys");

struct Build
{
    // Position-ordered, double-buffered: bounds_in[position] / perm_in[position] describe the primitive at that position when
    // the level starts (perm = its index in the caller's array); k_bounds<true> writes the permuted level into *_out and the host
    // swaps.  Every pass over positions therefore reads contiguously; only the elements a partition moves are gathered.
    const spcu_bounds* bounds_in;
    spcu_bounds*       bounds_out;
    const uint32_t*    perm_in;
    uint32_t*          perm_out;
    uint32_t           n;
    uint32_t*          seg_of; // [n] node (BFS numbering) the position belongs to at the current level; kInvalid once in a leaf
    // per node, BFS numbering, capacity 2n
    uint32_t* first;
    uint32_t* last;
    uint32_t* parent;   // parent << 1 | which child; kInvalid for the root
    uint32_t* left_run; // consecutive left-child steps that lead here (0 for the root and for right children)
    uint32_t* state;
    uint32_t* child0;   // children are allocated in pairs: child0, child0 + 1
    float*    box;      // [6] lo.xyz hi.xyz
    // per node of the current level (index = node - level_begin)
    NodeKeys* keys;
    uint32_t* dim;
    float*    at;
    uint32_t* mid;      // partition point of a node that became internal
    // per position
    uint8_t*  flag;     // partition predicate; reused for the chain counts of the numbering pass
    uint32_t* prefix;   // [n + 1] exclusive prefix sum of flag
    uint32_t* left_misplaced;
    uint32_t* right_misplaced;
    uint32_t* n_nodes;  // device counter
};

// ---- ordered keys ---------------------------------------------------------------------------------------------------
// float -> uint32, monotone, with -0 folded onto +0 so that the two tie (as they do under < and >).
__device__ __forceinline__ uint32_t ordered(float f)
{
    uint32_t u = __float_as_uint(f);
    if ((u << 1) == 0u) {
        u = 0u;
    }
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float unordered(uint32_t k, uint32_t negative_zero)
{
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float((u == 0u && negative_zero) ? 0x80000000u : u);
}

// smallest value; among equal values the LARGEST position; carries the sign bit of the winner
__device__ __forceinline__ unsigned long long min_key(float f, uint32_t pos)
{
    return (static_cast<unsigned long long>(ordered(f)) << 32) | ((0x7FFFFFFFu - pos) << 1) | (__float_as_uint(f) >> 31);
}
// largest value; among equal values the LARGEST position
__device__ __forceinline__ unsigned long long max_key(float f, uint32_t pos)
{
    return (static_cast<unsigned long long>(ordered(f)) << 32) | (pos << 1) | (__float_as_uint(f) >> 31);
}

__device__ __forceinline__ spcu_bounds load_bounds(const spcu_bounds* p)
{
    // 24-byte records: three 8-byte loads
    const float2* q = reinterpret_cast<const float2*>(p);
    const float2  a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    spcu_bounds   r;
    r.lo[0] = a.x, r.lo[1] = a.y, r.lo[2] = b.x, r.hi[0] = b.y, r.hi[1] = c.x, r.hi[2] = c.y;
    return r;
}

// ---- kernels --------------------------------------------------------------------------------------------------------
__global__ void k_init(Build b)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += stride) {
        b.perm_out[i] = i; // the host swaps in/out before the first level
        b.seg_of[i]   = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        b.first[0] = 0, b.last[0] = b.n, b.parent[0] = kInvalid, b.left_run[0] = 0, b.state[0] = kPending;
        *b.n_nodes = 1;
    }
}

__global__ void k_keys_init(NodeKeys* keys, uint32_t count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        NodeKeys k;
        k.k[0] = k.k[1] = k.k[2] = ~0ull;
        k.k[3] = k.k[4] = k.k[5] = 0ull;
        k.first_nonzero[0] = k.first_nonzero[1] = k.first_nonzero[2] = kInvalid;
        k.pad                                                        = 0;
        keys[i]                                                      = k;
    }
}

// Per-lane partial reduction of k_bounds: keys of the positions a lane has seen for node `seg`.
struct Partial
{
    unsigned long long k[6];
    uint32_t           fnz[3];
    uint32_t           seg;
};

// Positions of one node are contiguous: reduce each run of equal `seg` across the warp towards its first lane, which adds the
// run to the node's keys.  `start` = first position the lane's partial covers, `covered` = positions per lane-run element
// (32 * iterations for a partial accumulated over several iterations of a warp-uniform node, else 1 per lane).
__device__ __forceinline__ void flush_partial(const Build& b, uint32_t level_begin, Partial p, uint32_t start, uint32_t iterations)
{
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t other = __shfl_down_sync(0xFFFFFFFFu, p.seg, off);
        const bool     take  = lane + off < 32u && other == p.seg;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const unsigned long long lo = __shfl_down_sync(0xFFFFFFFFu, p.k[a], off);
            const unsigned long long hi = __shfl_down_sync(0xFFFFFFFFu, p.k[3 + a], off);
            const uint32_t           z  = __shfl_down_sync(0xFFFFFFFFu, p.fnz[a], off);
            if (take) {
                p.k[a]     = min(p.k[a], lo);
                p.k[3 + a] = max(p.k[3 + a], hi);
                p.fnz[a]   = min(p.fnz[a], z);
            }
        }
    }
    const uint32_t before = __shfl_up_sync(0xFFFFFFFFu, p.seg, 1);
    const bool     head   = lane == 0u || before != p.seg;
    const uint32_t heads  = __ballot_sync(0xFFFFFFFFu, head);
    if (head && p.seg != kInvalid) {
        const uint32_t above   = lane == 31u ? 0u : (heads >> (lane + 1u)) << (lane + 1u);
        const uint32_t run_len = ((above ? __ffs(above) - 1u : 32u) - lane) * iterations;
        NodeKeys*      dst     = &b.keys[p.seg - level_begin];
        if (start == b.first[p.seg] && start + run_len == b.last[p.seg]) {
            // the run is the whole node: single writer
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                dst->k[a] = p.k[a];
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                dst->first_nonzero[a] = p.fnz[a];
            }
        } else {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                atomicMin(&dst->k[a], p.k[a]);
                atomicMax(&dst->k[3 + a], p.k[3 + a]);
                if (p.fnz[a] != kInvalid) {
                    atomicMin(&dst->first_nonzero[a], p.fnz[a]);
                }
            }
        }
    }
}

// bounds = merge(bounds, prim->get_world_bounds()) over the node's range (BVHAccelerator.h:181-185), as key reductions.
// A warp walks kBoundsIterations x 32 consecutive positions.  While they all belong to one node (the upper levels of the
// tree, where every warp of the grid would otherwise hit the same nine addresses) the lanes just keep accumulating; the
// partial is flushed when the node changes.  Measured on 28 M boxes: the first levels took 5 ms each without this.
constexpr int kBoundsIterations = 32;

__device__ __forceinline__ void store_bounds(spcu_bounds* p, const spcu_bounds& x)
{
    float2* o = reinterpret_cast<float2*>(p);
    o[0] = make_float2(x.lo[0], x.lo[1]), o[1] = make_float2(x.lo[2], x.hi[0]), o[2] = make_float2(x.hi[1], x.hi[2]);
}

// kApply = false: the bounds of the level's nodes from the elements where they are (the root).
// kApply = true : FUSED with the end of the previous level — every position first fetches the element the partition leaves
// there (its own, or its partner's when it holds a misplaced one: Hoare's partition as a gather), writes it to the level's
// output buffers and joins its child, or retires with its leaf (whose elements stay put in BOTH buffers from then on); the
// element it has just moved then goes straight into the CHILD's keys.  One pass over the bounds instead of two.
// level_begin: first node of the level being partitioned; keys_begin: first node of the level whose keys are reduced.
template <bool kApply>
__global__ void __launch_bounds__(kBlock) k_bounds(Build b, uint32_t level_begin, uint32_t keys_begin)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t base = warp * (32u * kBoundsIterations);
    if (base >= b.n) {
        return;
    }
    Partial  carry;
    uint32_t carry_start = 0, carry_iterations = 0; // 0 iterations: nothing carried
    bool     carry_uniform = false;
    for (int it = 0; it < kBoundsIterations; ++it) {
        const uint64_t pos64 = base + static_cast<uint64_t>(it) * 32u + lane;
        if (pos64 - lane >= b.n) {
            break;
        }
        const uint32_t pos = static_cast<uint32_t>(pos64);
        Partial        cur;
        spcu_bounds    x{};
        cur.seg = kInvalid;
        const uint32_t seg = pos < b.n ? b.seg_of[pos] : kInvalid;
        if (seg != kInvalid) {
            if (!kApply) {
                cur.seg = seg;
                x       = load_bounds(&b.bounds_in[pos]);
            } else {
                uint32_t src = pos;
                if (b.state[seg] == kInternal) {
                    const uint32_t first = b.first[seg], mid = b.mid[seg - level_begin];
                    const uint32_t before = b.prefix[pos] - b.prefix[first]; // trues in [first, pos)
                    const bool     f      = b.flag[pos] != 0;
                    if (pos < mid && !f) {
                        src = b.right_misplaced[first + (pos - first - before)];
                    } else if (pos >= mid && f) {
                        src = b.left_misplaced[first + (mid - first - before - 1u)];
                    }
                    cur.seg = b.child0[seg] + (pos >= mid ? 1u : 0u);
                }
                b.seg_of[pos]   = cur.seg;
                b.perm_out[pos] = b.perm_in[src];
                x               = load_bounds(&b.bounds_in[src]);
                store_bounds(&b.bounds_out[pos], x);
            }
        }
        if (cur.seg != kInvalid) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                cur.k[a]     = min_key(x.lo[a], pos);
                cur.k[3 + a] = max_key(x.hi[a], pos);
                cur.fnz[a]   = (x.lo[a] == 0.0f && x.hi[a] == 0.0f) ? kInvalid : pos;
            }
        } else {
            cur.k[0] = cur.k[1] = cur.k[2] = ~0ull;
            cur.k[3] = cur.k[4] = cur.k[5] = 0ull;
            cur.fnz[0] = cur.fnz[1] = cur.fnz[2] = kInvalid;
        }
        const uint32_t seg0    = __shfl_sync(0xFFFFFFFFu, cur.seg, 0);
        const bool     uniform = __all_sync(0xFFFFFFFFu, cur.seg == seg0);
        if (carry_iterations && uniform && carry_uniform && seg0 == carry.seg) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                carry.k[a]     = min(carry.k[a], cur.k[a]);
                carry.k[3 + a] = max(carry.k[3 + a], cur.k[3 + a]);
                carry.fnz[a]   = min(carry.fnz[a], cur.fnz[a]);
            }
            ++carry_iterations;
            continue;
        }
        if (carry_iterations) {
            flush_partial(b, keys_begin, carry, carry_start, carry_uniform ? carry_iterations : 1u);
        }
        carry            = cur;
        carry_start      = pos;
        carry_iterations = 1;
        carry_uniform    = uniform;
    }
    if (carry_iterations) {
        flush_partial(b, keys_begin, carry, carry_start, carry_uniform ? carry_iterations : 1u);
    }
}

// Keys -> the bounds the reference's sequential fold ends with, bit for bit.  merge() returns BBox{min(lo, lo'), max(hi, hi')}
// (math/BBox.h:60-64) and that constructor sorts its two corners once more (:26-30): lo = min(l, h), hi = max(l, h), second
// operand on a tie.  The re-sort only shows while the running interval is [+-0, +-0]: then lo takes h's sign.  So: upper =
// the last maximum; lower = the last minimum x_j.lo, unless its value is zero and every primitive up to and including
// position j is the interval [+-0, +-0] on this axis, in which case it is x_j.hi.
__device__ void decode_bounds(const Build& b, const NodeKeys& nk, float lo[3], float hi[3])
{
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const unsigned long long kl = nk.k[a], kh = nk.k[3 + a];
        float                    l = unordered(static_cast<uint32_t>(kl >> 32), static_cast<uint32_t>(kl) & 1u);
        const float              h = unordered(static_cast<uint32_t>(kh >> 32), static_cast<uint32_t>(kh) & 1u);
        if (l == 0.0f) {
            const uint32_t j = 0x7FFFFFFFu - (static_cast<uint32_t>(kl) >> 1);
            if (j < nk.first_nonzero[a]) {
                l = b.bounds_in[j].hi[a];
            }
        }
        lo[a] = l;
        hi[a] = h;
    }
}

// max_dim (math/Vector3.h:653-670)
__device__ __forceinline__ int max_dim(float x, float y, float z)
{
    x = fabsf(x), y = fabsf(y), z = fabsf(z);
    if (x > y) {
        return x > z ? 0 : 2;
    }
    return y > z ? 1 : 2;
}

// Leaf or split (BVHAccelerator.h:187-195).
__global__ void k_decide(Build b, uint32_t level_begin, uint32_t level_count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= level_count) {
        return;
    }
    const uint32_t node = level_begin + i;
    float          lo[3], hi[3];
    decode_bounds(b, b.keys[i], lo, hi);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        b.box[6 * node + a]     = lo[a];
        b.box[6 * node + 3 + a] = hi[a];
    }
    if (b.last[node] - b.first[node] <= 4u) { // k_max_leaf_elements (:211)
        b.state[node] = kLeaf;
        return;
    }
    const int d   = max_dim(__fsub_rn(hi[0], lo[0]), __fsub_rn(hi[1], lo[1]), __fsub_rn(hi[2], lo[2]));
    b.dim[i]      = static_cast<uint32_t>(d);
    b.at[i]       = __fdiv_rn(__fadd_rn(lo[d], hi[d]), 2.0f); // center (math/BBox.h:114-118)
    b.state[node] = kSplitting;
}

// The partition predicate center(prim bounds)[d] < split (BVHAccelerator.h:196-198) for every position of a splitting node.
__global__ void __launch_bounds__(kBlock) k_flags(Build b, uint32_t level_begin)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < b.n; pos += stride) {
        const uint32_t seg = b.seg_of[pos];
        uint8_t        f   = 0;
        if (seg != kInvalid && b.state[seg] == kSplitting) {
            const uint32_t d = b.dim[seg - level_begin];
            const float*   x = reinterpret_cast<const float*>(&b.bounds_in[pos]);
            f                = __fdiv_rn(__fadd_rn(__ldg(x + d), __ldg(x + 3 + d)), 2.0f) < b.at[seg - level_begin];
        }
        b.flag[pos] = f;
    }
}

// std::partition's outcome for every splitting node: empty side -> leaf of any size (BVHAccelerator.h:200-203), else two children.
__global__ void k_split(Build b, uint32_t level_begin, uint32_t level_count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= level_count) {
        return;
    }
    const uint32_t node = level_begin + i;
    if (b.state[node] != kSplitting) {
        return;
    }
    const uint32_t first = b.first[node], last = b.last[node];
    const uint32_t trues = b.prefix[last] - b.prefix[first];
    if (trues == 0u || trues == last - first) {
        b.state[node] = kLeaf;
        return;
    }
    const uint32_t mid = first + trues;
    const uint32_t c0  = atomicAdd(b.n_nodes, 2u);
    b.state[node]      = kInternal;
    b.child0[node]     = c0;
    b.mid[i]           = mid;
    b.first[c0] = first, b.last[c0] = mid, b.parent[c0] = node << 1, b.left_run[c0] = b.left_run[node] + 1u, b.state[c0] = kPending;
    b.first[c0 + 1] = mid, b.last[c0 + 1] = last, b.parent[c0 + 1] = (node << 1) | 1u, b.left_run[c0 + 1] = 0u, b.state[c0 + 1] = kPending;
}

// Hoare partition as a permutation, step 1: list the misplaced elements of both sides in the order the two cursors of
// libstdc++'s __partition (bits/stl_algo.h, bidirectional version) meet them.
__global__ void __launch_bounds__(kBlock) k_scatter(Build b, uint32_t level_begin)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < b.n; pos += stride) {
        const uint32_t seg = b.seg_of[pos];
        if (seg == kInvalid || b.state[seg] != kInternal) {
            continue;
        }
        const uint32_t first = b.first[seg], mid = b.mid[seg - level_begin];
        const uint32_t before = b.prefix[pos] - b.prefix[first]; // trues in [first, pos)
        const bool     f      = b.flag[pos] != 0;
        if (pos < mid && !f) {
            b.left_misplaced[first + (pos - first - before)] = pos; // rank among the falses, from the left
        } else if (pos >= mid && f) {
            b.right_misplaced[first + (mid - first - before - 1u)] = pos; // trues after pos = rank from the right
        }
    }
}

// Numbering: internal nodes that start at the same position form a chain of left children; the chain's length goes to its
// start position, and the prefix sum over positions is the pre-order index of each chain's head.
__global__ void k_chain_counts(Build b, uint32_t n_nodes)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n_nodes && b.state[v] == kInternal && b.state[b.child0[v]] != kInternal) {
        b.flag[b.first[v]] = static_cast<uint8_t>(b.left_run[v] + 1u);
    }
}

// Every node but the root fills its half of its parent's record (the flattener's layout, include/spcu.h spcu_bvh_node).
__global__ void k_emit(Build b, uint32_t n_nodes, spcu_bvh_node* out, uint32_t first_id, const uint8_t* non_triangle)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v == 0u || v >= n_nodes) {
        return;
    }
    const uint32_t parent = b.parent[v] >> 1, which = b.parent[v] & 1u;
    spcu_bvh_node* dst    = &out[b.prefix[b.first[parent]] + b.left_run[parent]];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        dst->box[6 * which + a] = b.box[6 * v + a];
    }
    if (b.state[v] == kInternal) {
        dst->child[which] = static_cast<int32_t>(b.prefix[b.first[v]] + b.left_run[v]);
        dst->count[which] = 0u;
    } else {
        uint32_t mixed = 0;
        if (non_triangle) {
            for (uint32_t p = b.first[v]; p < b.last[v]; ++p) {
                mixed |= non_triangle[b.perm_in[p]];
            }
        }
        dst->child[which] = ~static_cast<int32_t>(first_id + b.first[v]);
        dst->count[which] = (b.last[v] - b.first[v]) | (mixed ? SPCU_LEAF_MIXED_FLAG : 0u);
    }
}

// Triangle::get_world_bounds_impl (shapes/Triangle.h:228-237): BBox::extend(p) = { min(p, m_min), max(p, m_max) }
// (math/BBox.h:42-46), _mm_min_ps / _mm_max_ps semantics (the running value is the second operand and wins a tie).
__global__ void __launch_bounds__(kBlock) k_triangle_bounds(const spcu_prim_geom* tris, uint32_t n, spcu_bounds* out)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4* q  = reinterpret_cast<const float4*>(&tris[i]);
        const float4  p0 = __ldg(q), p1 = __ldg(q + 1), p2 = __ldg(q + 2);
        const float   v[3][3] = { { p0.x, p0.y, p0.z }, { p1.x, p1.y, p1.z }, { p2.x, p2.y, p2.z } };
        float         lo[3], hi[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            lo[a] = __int_as_float(0x7F800000), hi[a] = __int_as_float(0xFF800000);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                lo[a] = v[k][a] < lo[a] ? v[k][a] : lo[a];
                hi[a] = v[k][a] > hi[a] ? v[k][a] : hi[a];
            }
        }
        float2* o = reinterpret_cast<float2*>(&out[i]);
        o[0]      = make_float2(lo[0], lo[1]);
        o[1]      = make_float2(lo[2], hi[0]);
        o[2]      = make_float2(hi[1], hi[2]);
    }
}

// Gathers the primitive records of the bounded primitives into leaf order: dst[first_id + k] = src[first_id + order[k]].
__global__ void __launch_bounds__(kBlock) k_gather_prims(const uint32_t* order, uint32_t n, uint32_t first_id, const float4* src_geom,
                                                         const float4* src_shade, const uint32_t* src_meta, float4* dst_geom,
                                                         float4* dst_shade, uint32_t* dst_meta)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const size_t from = first_id + order[k], to = first_id + k;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            dst_geom[3 * to + j]  = __ldg(&src_geom[3 * from + j]);
            dst_shade[3 * to + j] = __ldg(&src_shade[3 * from + j]);
        }
        dst_meta[to] = __ldg(&src_meta[from]);
    }
}

// Triangle bounds for the bounded primitives of a scene; non-triangles (spheres) keep what the caller supplied.
__global__ void __launch_bounds__(kBlock) k_scene_bounds(const spcu_prim_geom* prims, const uint32_t* meta, uint32_t n, spcu_bounds* out,
                                                         uint8_t* non_triangle)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const bool tri  = SPCU_META_KIND(meta[i]) == SPCU_PRIM_TRIANGLE;
        non_triangle[i] = tri ? 0 : 1;
        if (!tri) {
            continue;
        }
        const float4* q  = reinterpret_cast<const float4*>(&prims[i]);
        const float4  p0 = __ldg(q), p1 = __ldg(q + 1), p2 = __ldg(q + 2);
        const float   v[3][3] = { { p0.x, p0.y, p0.z }, { p1.x, p1.y, p1.z }, { p2.x, p2.y, p2.z } };
        spcu_bounds   r;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            r.lo[a] = __int_as_float(0x7F800000), r.hi[a] = __int_as_float(0xFF800000);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                r.lo[a] = v[k][a] < r.lo[a] ? v[k][a] : r.lo[a];
                r.hi[a] = v[k][a] > r.hi[a] ? v[k][a] : r.hi[a];
            }
        }
        out[i] = r;
    }
}

// The construction proper, on device-resident inputs.  Leaves the internal nodes in *d_nodes (allocated from `mem`), the
// leaf order in b.perm_in and fills `a` (a.nodes stays untouched).  d_non_tri may be NULL.
int device_build(spcu_ctx* c, Scratch& mem, Build& b, spcu_bounds* d_bounds, uint32_t n, const uint8_t* d_non_tri,
                 uint32_t first_id, uint32_t capacity, spcu_bvh_node** d_nodes, spcu_accel& a)
{
    const cudaStream_t st        = c->stream;
    const size_t       max_nodes = 2 * static_cast<size_t>(n);
    // a node with children holds >= 5 primitives, so a level has at most n/5 of them and the next level 2n/5 nodes
    const size_t       max_level = std::max<size_t>(2 * (static_cast<size_t>(n) / 5) + 2, 4);
    const uint32_t     n_tiles   = n / kScanTile + 1; // covers index n
    uint32_t*          partials  = nullptr;
    b   = Build{};
    b.n = n;
    spcu_bounds* bounds_buf[2] = { d_bounds, nullptr };
    uint32_t*    perm_buf[2]   = { nullptr, nullptr };
    CK(c, mem.get(&bounds_buf[1], n));
    CK(c, mem.get(&perm_buf[0], n));
    CK(c, mem.get(&perm_buf[1], n));
    int  cur     = 0; // buffers the current level reads
    auto set_buf = [&] {
        b.bounds_in = bounds_buf[cur], b.bounds_out = bounds_buf[cur ^ 1], b.perm_in = perm_buf[cur], b.perm_out = perm_buf[cur ^ 1];
    };
    cur = 1; // k_init writes perm_out = perm_buf[0]
    set_buf();
    CK(c, mem.get(&b.seg_of, n));
    CK(c, mem.get(&b.first, max_nodes));
    CK(c, mem.get(&b.last, max_nodes));
    CK(c, mem.get(&b.parent, max_nodes));
    CK(c, mem.get(&b.left_run, max_nodes));
    CK(c, mem.get(&b.state, max_nodes));
    CK(c, mem.get(&b.child0, max_nodes));
    CK(c, mem.get(&b.box, 6 * max_nodes));
    CK(c, mem.get(&b.keys, max_level));
    CK(c, mem.get(&b.dim, max_level));
    CK(c, mem.get(&b.at, max_level));
    CK(c, mem.get(&b.mid, max_level));
    CK(c, mem.get(&b.flag, static_cast<size_t>(n_tiles) * kScanTile));
    CK(c, mem.get(&b.prefix, static_cast<size_t>(n) + 1));
    CK(c, mem.get(&b.left_misplaced, n));
    CK(c, mem.get(&b.right_misplaced, n));
    CK(c, mem.get(&b.n_nodes, 1));
    CK(c, mem.get(&partials, n_tiles));
    CK(c, mem.get(d_nodes, std::min<size_t>(capacity, n - 1))); // a binary tree over n leaves at most; allocated before the timed kernels
    CK(c, cudaMemsetAsync(b.flag, 0, static_cast<size_t>(n_tiles) * kScanTile, st));

    const unsigned grid        = grid_for(n, c->sm_count);
    const unsigned bounds_grid = static_cast<unsigned>((static_cast<uint64_t>(n) + kBlock * kBoundsIterations - 1) / (kBlock * kBoundsIterations));
    CK(c, cudaEventRecord(c->ev0, st)); // ev0 .. ev1: the construction kernels (allocations above are outside)
    k_init<<<grid, kBlock, 0, st>>>(b);
    cur = 0;
    uint32_t level_begin = 0, level_end = 1, max_depth = 0;
    for (uint32_t level = 0; level_begin < level_end; ++level) {
        const uint32_t count = level_end - level_begin;
        if (count > max_level) {
            return fail(c, SPCU_ERR_INVALID, "spcu_build_bvh: level %u holds %u nodes (internal error)", level, count);
        }
        const unsigned node_grid = (count + kBlock - 1) / kBlock;
        set_buf();
        if (level == 0) {
            k_keys_init<<<1, kBlock, 0, st>>>(b.keys, 1);
            k_bounds<false><<<bounds_grid, kBlock, 0, st>>>(b, 0, 0);
        }
        k_decide<<<node_grid, kBlock, 0, st>>>(b, level_begin, count);
        k_flags<<<grid, kBlock, 0, st>>>(b, level_begin);
        exclusive_scan(b.flag, b.prefix, b.n, partials, st);
        k_split<<<node_grid, kBlock, 0, st>>>(b, level_begin, count);
        // the children's keys (k_decide has consumed this level's): at most two per node of this level
        const uint32_t next_max = static_cast<uint32_t>(std::min<size_t>(2 * static_cast<size_t>(count), max_level));
        k_keys_init<<<(next_max + kBlock - 1) / kBlock, kBlock, 0, st>>>(b.keys, next_max);
        k_scatter<<<grid, kBlock, 0, st>>>(b, level_begin);
        k_bounds<true><<<bounds_grid, kBlock, 0, st>>>(b, level_begin, level_end); // apply the partition + the children's bounds
        cur ^= 1;
        uint32_t total = 0;
        CK(c, cudaMemcpyAsync(&total, b.n_nodes, sizeof total, cudaMemcpyDeviceToHost, st));
        CK(c, cudaStreamSynchronize(st));
        if (total > level_end) {
            max_depth = level + 1;
            if (max_depth > SPCU_MAX_BVH_DEPTH) {
                return fail(c, SPCU_ERR_LIMIT, "spcu_build_bvh: tree deeper than SPCU_MAX_BVH_DEPTH (%u)", SPCU_MAX_BVH_DEPTH);
            }
        }
        level_begin = level_end;
        level_end   = total;
    }
    const uint32_t n_nodes = level_end;
    set_buf(); // perm_in = the final order

    // numbering + emission
    CK(c, cudaMemsetAsync(b.flag, 0, static_cast<size_t>(n_tiles) * kScanTile, st));
    const unsigned all_grid = (n_nodes + kBlock - 1) / kBlock;
    k_chain_counts<<<all_grid, kBlock, 0, st>>>(b, n_nodes);
    exclusive_scan(b.flag, b.prefix, b.n, partials, st);
    uint32_t n_internal = 0, root_state = 0;
    CK(c, cudaMemcpyAsync(&n_internal, b.prefix + n, sizeof n_internal, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(&root_state, b.state, sizeof root_state, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    if (n_internal > capacity) {
        return fail(c, SPCU_ERR_LIMIT, "spcu_build_bvh: %u internal nodes, capacity %u", n_internal, capacity);
    }
    k_emit<<<all_grid, kBlock, 0, st>>>(b, n_nodes, *d_nodes, first_id, d_non_tri);
    CK(c, cudaEventRecord(c->ev1, st));
    CK(c, cudaGetLastError());
    a.n_unbounded = first_id;
    a.n_prims     = first_id + n;
    a.n_nodes     = n_internal;
    a.max_depth   = max_depth;
    a.root_count  = 0;
    if (root_state == kInternal) {
        a.root = 0;
    } else { // one leaf holds everything: its mixed flag is the OR over all primitives
        uint32_t mixed = 0;
        if (d_non_tri) {
            std::vector<uint8_t> h(n);
            CK(c, cudaMemcpyAsync(h.data(), d_non_tri, n, cudaMemcpyDeviceToHost, st));
            CK(c, cudaStreamSynchronize(st));
            for (const uint8_t f : h) {
                mixed |= f;
            }
        }
        a.root       = ~static_cast<int32_t>(first_id);
        a.root_count = n | (mixed ? SPCU_LEAF_MIXED_FLAG : 0u);
    }
    return SPCU_OK;
}

} // namespace

// The geometry half of spcu_upload_scene_build (spcu_api.cu): primitive records arrive in PRE-construction order
// ([0, n_unbounded) the top-level list, then the bounded primitives as the reference holds them before
// BVHAccelerator(first, last)); bounds, tree and the gather into leaf order all happen on the device, and the context's
// geometry buffers end up exactly as spcu_upload_scene would have filled them from the flattener's output.
// ---- 4-wide nodes (trace.cuh): wide node i = the grandchildren of binary node i (a child that is a leaf stands for itself) ----
// Every child is described by ONE packed word (trace.cuh wide_unpack): a walk's stack entry is that word, so a pop finds the
// child's node or triangles without the dependent load of { link, count } round 2's first wide walk paid.  Leaves of more than
// kWideSmallLeafMax primitives (the reference makes them where std::partition cannot split: BVHAccelerator.h:195-198) do not
// fit the word and go through the side table `big` (one { first, count } per such leaf, slots handed out by `n_big`).
__global__ void k_build_wide(const float4* nodes, uint32_t n, float4* wide, int2* big, uint32_t* n_big)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float    box[4][6];
        int32_t  link[4];
        uint32_t count[4];
        int      m = 0;
        const float4 v0 = nodes[4 * i + 0], v1 = nodes[4 * i + 1], v2 = nodes[4 * i + 2], v3 = nodes[4 * i + 3];
        const float  own[2][6] = { { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y }, { v1.z, v1.w, v2.x, v2.y, v2.z, v2.w } };
        const int32_t  cl[2]   = { __float_as_int(v3.x), __float_as_int(v3.y) };
        const uint32_t cc[2]   = { __float_as_uint(v3.z), __float_as_uint(v3.w) };
        for (int k = 0; k < 2; ++k) {
            if (cl[k] < 0) { // a leaf child keeps its own box
                for (int d = 0; d < 6; ++d) box[m][d] = own[k][d];
                link[m]  = cl[k];
                count[m] = cc[k];
                ++m;
                continue;
            }
            const float4* c  = nodes + 4 * static_cast<size_t>(cl[k]);
            const float4  w0 = c[0], w1 = c[1], w2 = c[2], w3 = c[3];
            const float   g[2][6] = { { w0.x, w0.y, w0.z, w0.w, w1.x, w1.y }, { w1.z, w1.w, w2.x, w2.y, w2.z, w2.w } };
            for (int j = 0; j < 2; ++j) {
                for (int d = 0; d < 6; ++d) box[m][d] = g[j][d];
                link[m]  = j == 0 ? __float_as_int(w3.x) : __float_as_int(w3.y);
                count[m] = j == 0 ? __float_as_uint(w3.z) : __float_as_uint(w3.w);
                ++m;
            }
        }
        for (; m < 4; ++m) { // unused slot: a point box no ray of the NaN-free form enters, an empty leaf behind it
            for (int d = 0; d < 6; ++d) box[m][d] = 3.0e38f;
            link[m]  = ~0;
            count[m] = 0u;
        }
        // one 32-byte sector per child: { lo.xyz, hi.x } { hi.yz, packed child, count }
        float4* o = wide + 8 * static_cast<size_t>(i);
        for (int k = 0; k < 4; ++k) {
            int32_t packed = link[k]; // an internal node: its index
            if (link[k] < 0) {
                const uint32_t first = static_cast<uint32_t>(~link[k]), cnt = count[k] & SPCU_LEAF_COUNT_MASK;
                if (cnt <= kWideSmallLeafMax && (first <= kWidePayloadMask || cnt == 0u)) {
                    packed = static_cast<int32_t>(0x80000000u | ((count[k] & SPCU_LEAF_MIXED_FLAG) >> 1) | (cnt << 27) |
                                                  (cnt ? first : 0u));
                } else {
                    const uint32_t t = atomicAdd(n_big, 1u);
                    big[t]           = make_int2(link[k], static_cast<int32_t>(count[k]));
                    packed           = static_cast<int32_t>(0x80000000u | (kWideBigLeafTag << 27) | t);
                }
            }
            o[2 * k + 0] = make_float4(box[k][0], box[k][1], box[k][2], box[k][3]);
            o[2 * k + 1] = make_float4(box[k][4], box[k][5], __int_as_float(packed), __uint_as_float(count[k]));
        }
    }
}

void spcu::launch_build_wide(const float4* d_nodes, uint32_t n, float4* d_wide, int2* d_big, uint32_t* d_n_big, int sm_count,
                             cudaStream_t st)
{
    if (n) {
        cudaMemsetAsync(d_n_big, 0, sizeof(uint32_t), st);
        k_build_wide<<<grid_for(n, sm_count), kBlock, 0, st>>>(d_nodes, n, d_wide, d_big, d_n_big);
    }
}

// 0 into *flag when a primitive's bounds are not finite with lo <= hi (then the tree's boxes are not either)
__global__ void k_bounds_proper(const spcu_bounds* bounds, uint32_t n, int* flag)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const spcu_bounds b  = bounds[i];
        bool              ok = true;
        for (int d = 0; d < 3; ++d) {
            ok = ok && fabsf(b.lo[d]) < __int_as_float(0x7f800000) && fabsf(b.hi[d]) < __int_as_float(0x7f800000) && b.lo[d] <= b.hi[d];
        }
        if (!ok) {
            *flag = 0;
        }
    }
}

int spcu::build_scene_geometry(spcu_ctx* c, const spcu_flat_scene* s, const spcu_bounds* extra_bounds, uint32_t* order_out,
                               spcu_accel* built, bool* proper_boxes)
{
    *proper_boxes = true;
    const uint32_t n_prims = s->geom.n_prims, first_id = s->geom.n_unbounded, n = n_prims - first_id;
    const cudaStream_t st  = c->stream;
    Scratch            mem;
    spcu_prim_geom*    src_geom  = nullptr;
    spcu_prim_shade*   src_shade = nullptr;
    uint32_t*          src_meta  = nullptr;
    const size_t       np        = std::max<size_t>(n_prims, 1);
    CK(c, mem.get(&src_geom, np));
    CK(c, mem.get(&src_shade, np));
    CK(c, mem.get(&src_meta, np));
    CK(c, c->geom_prims.reserve(np * sizeof(spcu_prim_geom)));
    CK(c, c->geom_shade.reserve(np * sizeof(spcu_prim_shade)));
    CK(c, c->geom_meta.reserve(np * sizeof(uint32_t)));
    if (n_prims) {
        int rc;
        if ((rc = copy_to_device(c, src_geom, s->geom_prims, n_prims * sizeof(spcu_prim_geom))) != SPCU_OK) return rc;
        if ((rc = copy_to_device(c, src_shade, s->geom_shade, n_prims * sizeof(spcu_prim_shade))) != SPCU_OK) return rc;
        if ((rc = copy_to_device(c, src_meta, s->geom_meta, n_prims * sizeof(uint32_t))) != SPCU_OK) return rc;
        c->scene_bytes += n_prims * (sizeof(spcu_prim_geom) + sizeof(spcu_prim_shade) + sizeof(uint32_t));
    }
    if (first_id) { // the top-level list keeps its order
        CK(c, cudaMemcpyAsync(c->geom_prims.p, src_geom, first_id * sizeof(spcu_prim_geom), cudaMemcpyDeviceToDevice, st));
        CK(c, cudaMemcpyAsync(c->geom_shade.p, src_shade, first_id * sizeof(spcu_prim_shade), cudaMemcpyDeviceToDevice, st));
        CK(c, cudaMemcpyAsync(c->geom_meta.p, src_meta, first_id * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    spcu_accel a{};
    a.n_unbounded = first_id;
    a.n_prims     = n_prims;
    a.root        = ~static_cast<int32_t>(first_id);
    CK(c, c->geom_nodes.reserve(16));
    if (n) {
        spcu_bounds* d_bounds  = nullptr;
        uint8_t*     d_non_tri = nullptr;
        CK(c, mem.get(&d_bounds, n));
        CK(c, mem.get(&d_non_tri, n));
        if (extra_bounds) {
            CK(c, cudaMemcpyAsync(d_bounds, extra_bounds, static_cast<size_t>(n) * sizeof(spcu_bounds), cudaMemcpyHostToDevice, st));
        }
        k_scene_bounds<<<grid_for(n, c->sm_count), kBlock, 0, st>>>(src_geom + first_id, src_meta + first_id, n, d_bounds, d_non_tri);
        int* d_proper = nullptr;
        int  h_proper = 1;
        CK(c, mem.get(&d_proper, 1));
        CK(c, cudaMemcpyAsync(d_proper, &h_proper, sizeof(int), cudaMemcpyHostToDevice, st));
        k_bounds_proper<<<grid_for(n, c->sm_count), kBlock, 0, st>>>(d_bounds, n, d_proper);
        CK(c, cudaMemcpyAsync(&h_proper, d_proper, sizeof(int), cudaMemcpyDeviceToHost, st));
        Build          b;
        spcu_bvh_node* d_nodes = nullptr;
        if (const int rc = device_build(c, mem, b, d_bounds, n, d_non_tri, first_id, n - 1, &d_nodes, a); rc != SPCU_OK) {
            return rc;
        }
        CK(c, c->geom_nodes.reserve(std::max<size_t>(a.n_nodes, 1) * sizeof(spcu_bvh_node)));
        if (a.n_nodes) {
            CK(c, cudaMemcpyAsync(c->geom_nodes.p, d_nodes, a.n_nodes * sizeof(spcu_bvh_node), cudaMemcpyDeviceToDevice, st));
        }
        k_gather_prims<<<grid_for(n, c->sm_count), kBlock, 0, st>>>(
            b.perm_in, n, first_id, reinterpret_cast<const float4*>(src_geom), reinterpret_cast<const float4*>(src_shade), src_meta,
            c->geom_prims.as<float4>(), c->geom_shade.as<float4>(), c->geom_meta.as<uint32_t>());
        CK(c, cudaGetLastError());
        if (order_out) {
            CK(c, cudaMemcpyAsync(order_out, b.perm_in, static_cast<size_t>(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        }
        CK(c, cudaStreamSynchronize(st));
        *proper_boxes = h_proper != 0;
    }
    *built = a;
    return SPCU_OK;
}

extern "C" int spcu_build_bvh(spcu_ctx* c, const spcu_bounds* bounds, uint32_t n, const uint8_t* non_triangle, uint32_t first_id,
                              uint32_t* order, spcu_bvh_node* nodes, uint32_t capacity, spcu_accel* accel, float* device_ms)
{
    if (!c) {
        return SPCU_ERR_INVALID;
    }
    if (!accel || (n && (!bounds || !order)) || (capacity && !nodes)) {
        return fail(c, SPCU_ERR_INVALID, "spcu_build_bvh: NULL argument");
    }
    if (n >= (1u << 30) || static_cast<uint64_t>(first_id) + n >= (1u << 30)) {
        return fail(c, SPCU_ERR_LIMIT, "spcu_build_bvh: more than 2^30 primitives");
    }
    CK(c, cudaSetDevice(c->device));
    spcu_accel a{};
    a.n_unbounded = first_id;
    a.n_prims     = first_id + n;
    a.nodes       = nodes;
    if (device_ms) {
        *device_ms = 0.0f;
    }
    if (n == 0) { // BVHAccelerator(first, first): one empty leaf (base/Scene.h:34-37)
        a.root = ~static_cast<int32_t>(first_id);
        *accel = a;
        return SPCU_OK;
    }
    const cudaStream_t st = c->stream;
    Scratch            mem;
    spcu_bounds*       d_bounds  = nullptr;
    uint8_t*           d_non_tri = nullptr;
    CK(c, mem.get(&d_bounds, n));
    if (const int rc = copy_to_device(c, d_bounds, bounds, static_cast<size_t>(n) * sizeof(spcu_bounds)); rc != SPCU_OK) {
        return rc;
    }
    if (non_triangle) {
        CK(c, mem.get(&d_non_tri, n));
        CK(c, cudaMemcpyAsync(d_non_tri, non_triangle, n, cudaMemcpyHostToDevice, st));
    }
    Build          b;
    spcu_bvh_node* d_nodes = nullptr;
    if (const int rc = device_build(c, mem, b, d_bounds, n, d_non_tri, first_id, capacity, &d_nodes, a); rc != SPCU_OK) {
        return rc;
    }
    if (a.n_nodes) {
        CK(c, cudaMemcpyAsync(nodes, d_nodes, static_cast<size_t>(a.n_nodes) * sizeof(spcu_bvh_node), cudaMemcpyDeviceToHost, st));
    }
    CK(c, cudaMemcpyAsync(order, b.perm_in, static_cast<size_t>(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    if (device_ms) {
        CK(c, cudaEventElapsedTime(device_ms, c->ev0, c->ev1));
    }
    *accel = a;
    return SPCU_OK;
}

extern "C" int spcu_triangle_bounds(spcu_ctx* c, const spcu_prim_geom* tris, uint32_t n, spcu_bounds* out)
{
    if (!c) {
        return SPCU_ERR_INVALID;
    }
    if (n && (!tris || !out)) {
        return fail(c, SPCU_ERR_INVALID, "spcu_triangle_bounds: NULL argument");
    }
    CK(c, cudaSetDevice(c->device));
    if (n == 0) {
        return SPCU_OK;
    }
    Scratch         mem;
    spcu_prim_geom* d_tris = nullptr;
    spcu_bounds*    d_out  = nullptr;
    CK(c, mem.get(&d_tris, n));
    CK(c, mem.get(&d_out, n));
    CK(c, cudaMemcpyAsync(d_tris, tris, static_cast<size_t>(n) * sizeof(spcu_prim_geom), cudaMemcpyHostToDevice, c->stream));
    k_triangle_bounds<<<grid_for(n, c->sm_count), kBlock, 0, c->stream>>>(d_tris, n, d_out);
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(out, d_out, static_cast<size_t>(n) * sizeof(spcu_bounds), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return SPCU_OK;
}
