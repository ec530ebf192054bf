// Cave generator shared by caveflyer and jumper — device restatement of Room_Generator
// (games/caveflyer/room_generator.cpp == games/jumper/room_generator.cpp, byte-identical):
//   update()          :21-36   cellular automaton, Moore neighbourhood incl. self, out of bounds = wall
//   build_room()      :38-78   BFS flood fill into an unordered_set (start cell only enters when re-discovered, Q19)
//   find_best_room()  :143-164 first strictly largest room; its unordered_set ITERATION ORDER is observable,
//                              because the callers turn it into the free_cells vector (caveflyer/tilemap.cpp:158-161,
//                              jumper/tilemap.cpp:150-153) — emulated with pg2::USet
//   find_path()       :80-141  BFS with parent links, `covered` omits the source (Q19)
//   expand_room()     :166-207 4 rounds of 8-neighbour dilation restricted to space cells (order-insensitive)
// Bulk passes (automaton, dilation) are spread over the 32 lanes of the warp; the order-sensitive
// searches run on lane 0 and publish their result through shared memory.
#pragma once
#include "pg2_uset.cuh"
#include "pg2_warp.cuh"

namespace pg2 {

constexpr int ROOM_MAX_DIM = 45;         // world_dim: 20 easy, 40 hard, 45 memory (bit rows: <= 64)
constexpr int ROOM_MAX_BUCKETS = 2400;   // bucket count of a fresh unordered_set<int> holding <= 2025 keys: <= 2357; also >= 45 * 45 cell claims

// Atomically claim a byte flag (0 -> 1); true for the one thread that claimed it.
PG2_DEV bool claim_cell(uint8_t* p) {
#ifdef PG2_HOSTSIM
    if (*p) return false;
    *p = 1;
    return true;
#else
    uint32_t* word = (uint32_t*)((size_t)p & ~(size_t)3);
    uint32_t bit = 1u << (8u * (uint32_t)((size_t)p & 3));
    return (atomicOr(word, bit) & (0xffu << (8u * (uint32_t)((size_t)p & 3)))) == 0u;
#endif
}

// ITERATION ORDER of a fresh std::unordered_set<int> (libstdc++ 13, see pg2_uset.cuh) after inserting the distinct
// keys[0 .. n) in that order — computed by the whole warp instead of replaying n insertions.
//   Both _M_insert_bucket_begin and _M_rehash_aux(unique) place a node at the BEGINNING of its bucket's group, or — for
//   a bucket seen for the first time — at the very front of the list. Folding that rule over a sequence Q from an empty
//   table with nb buckets therefore yields
//       G(Q, nb) = groups by bucket (key % nb), groups in REVERSE order of first appearance in Q, each group in
//                  REVERSE order of appearance,
//   a rehash to nb' turns the list L into G(L, nb'), and the keys inserted until the next rehash extend that fold:
//       L_i = G(L_{i-1} ++ batch_i, nb_i).
//   For a fresh set every policy trigger is a rehash (load factor 1.0: 13, 29, 59, 127, 257, 541, 1109, 2357 buckets),
//   so <= 9 regroupings, each a counting sort over positions (first position and size per bucket, a descending scan
//   for the group offsets, a descending placement pass with in-warp ranking of equal buckets).
// bufa / bufb: >= n keys each; first / cnt: >= ROOM_MAX_BUCKETS ints each. Returns the buffer holding the order.
struct USetOrder {
    PG2_DEV static int mod_by(int key, int n, uint32_t magic) {
        if (n == 1) return 0;
        uint32_t q = (uint32_t)(((uint64_t)(uint32_t)key * magic) >> 32);
        int r = key - (int)q * n;
        return r >= n ? r - n : (r < 0 ? r + n : r);
    }

    // out = G(q[0 .. n), nb)
    PG2_DEV_NOINLINE static void regroup(WarpCtx& w, const uint16_t* q, int n, int nb, int* first, int* cnt, uint16_t* out) {
        const int lane = w.lane;
        const uint32_t magic = (uint32_t)(0xffffffffu / (uint32_t)nb) + 1u;
        __syncwarp();
        for (int b = lane; b < nb; b += WARP_LANES) { first[b] = 0x7fffffff; cnt[b] = 0; }
        __syncwarp();
        for (int j = lane; j < n; j += WARP_LANES) {
            const int b = mod_by(q[j], nb, magic);
            atomicMin(&first[b], j);
            atomicAdd(&cnt[b], 1);
        }
        __syncwarp();
        // group offsets: walk the positions downwards; a bucket's group starts where the groups of all buckets that
        // first appear later end. Only the lane holding a bucket's first position touches cnt[b]: size -> offset in place.
        int run = 0;
        for (int base = n - 1; base >= 0; base -= WARP_LANES) {
            const int j = base - lane;
            const int b = j >= 0 ? mod_by(q[j], nb, magic) : 0;
            const bool is_first = j >= 0 && first[b] == j;
            int total;
            const int excl = warp_excl_scan(is_first ? cnt[b] : 0, &total);
            if (is_first) cnt[b] = run + excl;
            run += total;
        }
        __syncwarp();
        // placement, positions downwards again: cnt[b] = next free slot of the group
        for (int base = n - 1; base >= 0; base -= WARP_LANES) {
            const int j = base - lane;
            const int b = j >= 0 ? mod_by(q[j], nb, magic) : 0;
            const uint32_t same = match_lanes(j >= 0 ? (uint32_t)b : 0x80000000u | (uint32_t)lane);
            const int rank = __popc(same & ((1u << lane) - 1u));   // lower lanes = later positions of the same bucket
            const int slot = j >= 0 ? cnt[b] : 0;
            __syncwarp();
            if (j >= 0) {
                out[slot + rank] = q[j];
                if ((same >> lane) == 1u) cnt[b] = slot + rank + 1;   // the last lane of the bucket in this chunk
            }
            __syncwarp();
        }
    }

    PG2_DEV_NOINLINE static uint16_t* order(WarpCtx& w, const uint16_t* keys, int n, uint16_t* bufa, uint16_t* bufb, int* first, int* cnt) {
        uint16_t* cur = bufa;   // L_{i-1}, extended in place by the keys of the running batch
        uint16_t* other = bufb;
        int nb = 1, next_resize = 0, count = 0, listed = 0;   // listed = length of the grouped prefix of cur
        int i = 0;
        while (i < n) {
            if (count + 1 > next_resize) {   // _Prime_rehash_policy::_M_need_rehash (load factor 1.0)
                const int floor_bkts = next_resize ? 0 : 11;
                const int min_bkts = count + 1 > floor_bkts ? count + 1 : floor_bkts;
                if (min_bkts >= nb) {
                    const int want = min_bkts + 1 > nb * 2 ? min_bkts + 1 : nb * 2;
                    if (count > listed) {    // close the running batch under the old bucket count
                        regroup(w, cur, count, nb, first, cnt, other);
                        uint16_t* t = cur; cur = other; other = t;
                    }
                    listed = count;
                    nb = USet<4, 4>::next_bkt(want, &next_resize);
                } else {
                    next_resize = nb;
                }
            }
            int take = next_resize - count;
            if (take > n - i) take = n - i;
            __syncwarp();
            for (int t = w.lane; t < take; t += WARP_LANES) cur[count + t] = keys[i + t];
            __syncwarp();
            i += take; count += take;
        }
        if (count > 0) {
            regroup(w, cur, count, nb, first, cnt, other);
            cur = other;
        }
        return cur;
    }
};

struct RoomGen {
    int W, H;
    uint8_t* grid;        // [y + H * x]: 1 wall, 0 space
    uint8_t* tmp;         // double buffer / membership scratch
    uint8_t* mark;        // visited / in-room / covered
    uint16_t* queue;      // BFS queue, `expanded` of find_path
    uint16_t* parents;
    uint16_t* order;      // best_room in unordered_set iteration order
    uint16_t* path;       // goal_path (src .. dst)
    uint16_t* seq;        // second key buffer of USetOrder
    int* res;             // lane 0 -> all lanes: [0] best_room size, [1] path length, [2] BFS tail, [3] queue index of dst
    int* claim;           // per cell: smallest (lane * 4 + neighbour) that wants it in the running BFS chunk
    int* scratch;         // ROOM_MAX_BUCKETS ints (USetOrder)
    uint64_t* rows_a;     // bit rows (bit y of rows[x] = cell (x, y)): grid / member / scratch of update() and expand()
    uint64_t* rows_b;
    uint64_t* rows_c;

    PG2_DEV_NOINLINE void init(WarpCtx& w, int width, int height) {
        W = width; H = height;
        const int cells = (W * H + 3) & ~3;          // scratch sized by THIS world (20 x 20 ... 45 x 45)
        grid = w.alloc<uint8_t>(cells);
        tmp = w.alloc<uint8_t>(cells);
        mark = w.alloc<uint8_t>(cells);
        queue = w.alloc<uint16_t>(cells + 64);
        parents = w.alloc<uint16_t>(cells + 64);
        order = w.alloc<uint16_t>(cells + 64);
        path = w.alloc<uint16_t>(cells + 64);
        seq = w.alloc<uint16_t>(cells + 64);
        res = w.alloc<int>(4);
        claim = w.alloc<int>(ROOM_MAX_BUCKETS);     // W * H claims during a BFS, bucket table afterwards
        scratch = w.alloc<int>(ROOM_MAX_BUCKETS);
        rows_a = w.alloc<uint64_t>(W); rows_b = w.alloc<uint64_t>(W); rows_c = w.alloc<uint64_t>(W);
    }

    PG2_DEV int get(int x, int y) const { return (x < 0 || y < 0 || x >= W || y >= H) ? 1 : grid[y + H * x]; }

    // ---- bit rows: rows[x] bit y = bytes[y + H * x] (0 / 1 bytes; H <= 64). H a multiple of 4 (20, 40): four cells per
    // 32-bit word; otherwise (45) byte by byte.
    PG2_DEV void bytes_to_rows(WarpCtx& w, const uint8_t* bytes, uint64_t* rows) {
        for (int x = w.lane; x < W; x += WARP_LANES) {
            uint64_t m = 0;
            if (H % 4 == 0) {
                const uint32_t* p = (const uint32_t*)(bytes + H * x);
                for (int k = 0; k < H / 4; k++)   // bytes b0..b3 of a word -> bits 21..24 of word * 0x204081
                    m |= (uint64_t)(((p[k] & 0x01010101u) * 0x00204081u >> 21) & 15u) << (4 * k);
            } else {
                for (int y = 0; y < H; y++) m |= (uint64_t)(bytes[y + H * x] & 1u) << y;
            }
            rows[x] = m;
        }
        __syncwarp();
    }
    PG2_DEV void rows_to_bytes(WarpCtx& w, const uint64_t* rows, uint8_t* bytes) {
        for (int x = w.lane; x < W; x += WARP_LANES) {
            const uint64_t m = rows[x];
            if (H % 4 == 0) {
                uint32_t* p = (uint32_t*)(bytes + H * x);
                for (int k = 0; k < H / 4; k++) p[k] = ((uint32_t)(m >> (4 * k)) & 15u) * 0x00204081u & 0x01010101u;
            } else {
                for (int y = 0; y < H; y++) bytes[y + H * x] = (uint8_t)(m >> y & 1ull);
            }
        }
        __syncwarp();
    }

    // Room_Generator::update: a cell becomes wall iff >= 5 of the 9 cells of its Moore neighbourhood (itself included,
    // out of bounds = wall) are walls. One lane per grid column x on 64-bit rows: the nine neighbour rows are added
    // bit-sliced (carry-save adders), count >= 5 <=> eights | (fours & (twos | ones)).
    PG2_DEV_NOINLINE void update(WarpCtx& w) {
        __syncwarp();
        bytes_to_rows(w, grid, rows_a);
        const uint64_t ALL = H >= 64 ? ~0ull : (1ull << H) - 1ull;
        for (int x = w.lane; x < W; x += WARP_LANES) {
            const uint64_t r[3] = { x > 0 ? rows_a[x - 1] : ALL, rows_a[x], x < W - 1 ? rows_a[x + 1] : ALL };
            uint64_t ones[3], twos[3];
            for (int k = 0; k < 3; k++) {
                const uint64_t up = ((r[k] << 1) | 1ull) & ALL;             // neighbour y - 1 (y = 0: out of bounds = wall)
                const uint64_t dn = (r[k] >> 1) | (1ull << (H - 1));        // neighbour y + 1 (y = H-1: out of bounds)
                ones[k] = r[k] ^ up ^ dn;
                twos[k] = (r[k] & up) | (dn & (r[k] ^ up));
            }
            const uint64_t s1 = ones[0] ^ ones[1] ^ ones[2];                                   // weight 1
            const uint64_t c1 = (ones[0] & ones[1]) | (ones[2] & (ones[0] ^ ones[1]));         // weight 2
            const uint64_t t = twos[0] ^ twos[1] ^ twos[2];                                    // weight 2
            const uint64_t c2 = (twos[0] & twos[1]) | (twos[2] & (twos[0] ^ twos[1]));         // weight 4
            const uint64_t s2 = t ^ c1, c3 = t & c1;                                           // weight 2, weight 4
            const uint64_t s4 = c2 ^ c3, s8 = c2 & c3;                                         // weight 4, weight 8
            rows_b[x] = (s8 | (s4 & (s2 | s1))) & ALL;
        }
        __syncwarp();
        rows_to_bytes(w, rows_b, grid);
    }

    // The queue discipline of build_room / find_path (room_generator.cpp:38-78 / 80-141) by the whole warp:
    // queue[0] = src, NOT marked (it re-enters the queue when a neighbour re-discovers it, Q19); 32 queued cells are
    // expanded per round, a discovered cell goes to the expansion that comes first in the serial order (queue
    // position, then neighbour order (x-1,y) (x,y-1) (x,y+1) (x+1,y)) through an atomicMin on claim[], and the
    // winners are appended in exactly that order (prefix sum over the lanes' winner counts).
    // Precondition: mark[] all 0. dst >= 0: stop once dst is queued (its parent chain is final), res[3] = its index.
    // Returns the queue length; parents[] (if wanted) = queue index of the discoverer.
    PG2_DEV_NOINLINE int bfs_ordered(WarpCtx& w, int src, int dst, bool with_parents) {
        const int lane = w.lane, cells = W * H;
        __syncwarp();
        for (int i = lane; i < cells; i += WARP_LANES) claim[i] = 0x7fffffff;
        if (lane == 0) { queue[0] = (uint16_t)src; parents[0] = 0xffff; res[3] = -1; }
        __syncwarp();
        int head = 0, tail = 1;
        while (head < tail) {
            const int chunk = tail - head < WARP_LANES ? tail - head : WARP_LANES;
            const int q = head + lane;
            const bool valid = lane < chunk;
            const int cur = valid ? queue[q] : 0, x = cur / H, y = cur % H;
            const int nx[4] = { x - 1, x, x, x + 1 }, ny[4] = { y, y - 1, y + 1, y };
            int cell[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                cell[k] = -1;
                if (valid && nx[k] >= 0 && ny[k] >= 0 && nx[k] < W && ny[k] < H) {
                    const int nxt = ny[k] + H * nx[k];
                    if (!mark[nxt] && grid[nxt] == 0) { cell[k] = nxt; atomicMin(&claim[nxt], lane * 4 + k); }
                }
            }
            __syncwarp();
            int wins = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (cell[k] >= 0 && claim[cell[k]] != lane * 4 + k) cell[k] = -1;
                wins += cell[k] >= 0;
            }
            int total;
            int pos = tail + warp_excl_scan(wins, &total);
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (cell[k] >= 0) {
                    queue[pos] = (uint16_t)cell[k];
                    if (with_parents) parents[pos] = (uint16_t)q;
                    mark[cell[k]] = 1;
                    if (cell[k] == dst) res[3] = pos;
                    pos++;
                }
            __syncwarp();
            head += chunk; tail += total;
            if (dst >= 0 && res[3] >= 0) break;
        }
        __syncwarp();
        return tail;
    }

    // Room_Generator::find_best_room -> order[0 .. n) = iteration order of best_room; returns n
    PG2_DEV_NOINLINE int find_best_room(WarpCtx& w) {
        // pass 1 (order-insensitive): size of every room in scan order. A one-cell room yields an EMPTY set (its
        // start cell is never re-discovered); the first room always replaces the initial best_room_size of -1.
        const int cells = W * H, lane = w.lane;
        __syncwarp();
        for (int i = lane; i < cells; i += WARP_LANES) mark[i] = 0;
        __syncwarp();
        int best_start = -1, best_size = -1;
        for (int base = 0; base < cells; base += WARP_LANES) {
            for (;;) {   // next unvisited space cell of this chunk, in scan order
                const int i = base + lane;
                const uint32_t m = lane_ballot(i < cells && grid[i] == 0 && !mark[i], lane);
                if (!m) break;
                const int start = base + __ffs(m) - 1;
                __syncwarp();
                if (lane == 0) { mark[start] = 1; queue[0] = (uint16_t)start; res[2] = 1; }
                __syncwarp();
                int head = 0, tail = 1;
                while (head < tail) {                            // one BFS level chunk per iteration
                    for (int q = head + lane; q < tail; q += WARP_LANES) {
                        int cur = queue[q], x = cur / H, y = cur % H;
                        const int nx[4] = { x - 1, x, x, x + 1 }, ny[4] = { y, y - 1, y + 1, y };
                        for (int k = 0; k < 4; k++) {
                            if (nx[k] < 0 || ny[k] < 0 || nx[k] >= W || ny[k] >= H) continue;
                            int nxt = ny[k] + H * nx[k];
                            if (grid[nxt] == 0 && claim_cell(&mark[nxt])) queue[atomicAdd(&res[2], 1)] = (uint16_t)nxt;
                        }
                    }
                    __syncwarp();
                    head = tail; tail = res[2];
                    __syncwarp();
                }
                int size = tail >= 2 ? tail : 0;
                if (size > best_size) { best_size = size; best_start = start; }
            }
        }
        __syncwarp();
        for (int i = lane; i < cells; i += WARP_LANES) mark[i] = 0;
        __syncwarp();
        // pass 2 (ordered): the winning room again, in the order build_room inserts it into the unordered_set
        int n = 0;
        if (best_size > 0) {
            const int tail = bfs_ordered(w, best_start, -1, false);
            n = tail - 1;                                        // every queued cell but the initial queue[0] was inserted
            uint16_t* o = USetOrder::order(w, queue + 1, n, order, seq, claim, scratch);
            __syncwarp();
            if (o != order) for (int i = lane; i < n; i += WARP_LANES) order[i] = o[i];
            __syncwarp();
        }
        return n;
    }

    // Room_Generator::find_path -> path[0 .. len) from src to dst; returns len (0: none)
    PG2_DEV_NOINLINE int find_path(WarpCtx& w, int src, int dst) {
        const int lane = w.lane;
        __syncwarp();
        for (int i = lane; i < W * H; i += WARP_LANES) mark[i] = 0;   // `covered` (does not contain src)
        __syncwarp();
        if (grid[src] != 0) return 0;
        int at = 0;                                                  // queue index of dst
        if (src != dst) {
            bfs_ordered(w, src, dst, true);
            at = res[3];
            if (at < 0) return 0;
        }
        __syncwarp();
        if (lane == 0) {
            int len = 0, k = at;
            if (src == dst) { queue[0] = (uint16_t)src; parents[0] = 0xffff; }
            while (k != 0xffff) { len++; k = parents[k]; }
            k = at;
            for (int j = len - 1; j >= 0; j--) { path[j] = queue[k]; k = parents[k]; }
            res[1] = len;
        }
        __syncwarp();
        return res[1];
    }

    // wide_path = goal_path dilated `rounds` times (Room_Generator::expand_room); member[] is the result. A space cell
    // joins when one of its 8 neighbours is a member on a space cell: on bit rows that is three rows OR-ed with their
    // one-bit shifts, per round and lane (= grid column).
    PG2_DEV_NOINLINE void expand(WarpCtx& w, const uint16_t* cells_in, int n, int rounds, uint8_t* member) {
        __syncwarp();
        for (int i = w.lane; i < W * H; i += WARP_LANES) member[i] = 0;
        __syncwarp();
        for (int i = w.lane; i < n; i += WARP_LANES) member[cells_in[i]] = 1;
        __syncwarp();
        bytes_to_rows(w, grid, rows_a);       // walls
        bytes_to_rows(w, member, rows_b);     // members
        const uint64_t ALL = H >= 64 ? ~0ull : (1ull << H) - 1ull;
        for (int r = 0; r < rounds; r++) {
            for (int x = w.lane; x < W; x += WARP_LANES) rows_c[x] = rows_b[x] & ~rows_a[x];   // members on space cells
            __syncwarp();
            for (int x = w.lane; x < W; x += WARP_LANES) {
                const uint64_t qa = x > 0 ? rows_c[x - 1] : 0ull, qb = rows_c[x], qc = x < W - 1 ? rows_c[x + 1] : 0ull;
                const uint64_t near = qa | (qa << 1) | (qa >> 1) | (qb << 1) | (qb >> 1) | qc | (qc << 1) | (qc >> 1);
                rows_b[x] |= near & ~rows_a[x] & ALL;
            }
            __syncwarp();
        }
        rows_to_bytes(w, rows_b, member);
    }
};

}  // namespace pg2
