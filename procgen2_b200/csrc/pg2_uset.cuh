// Emulation of the ITERATION ORDER of libstdc++ 13's std::unordered_set<int> (identity hash,
// unique keys). The reference walks `System::entities` (ecs.h:212-215) — and a few generator-
// local sets — in places where the order is observable: sprite draw order, RNG consumption per
// mob, "first candidate" picks. Restated from
//   bits/hashtable.h         _M_insert_unique_node / _M_insert_bucket_begin (2009-2035),
//                            _M_rehash_aux(unique) (2584-2618), _M_erase / _M_remove_bucket_begin
//   bits/hashtable_policy.h  _Prime_rehash_policy (674-700)
//   src/c++11/hashtable_c++0x.cc  _M_need_rehash, _M_next_bkt, __prime_list
// clear() keeps the bucket array, so the bucket count of an ECS system set is monotone over the
// lifetime of a process: it is per-env persistent state (`nb` in / out).
#pragma once
#include "pg2_platform.cuh"

namespace pg2 {

template <int MAXK, int MAXB>
struct USet {
    static constexpr int16_t END = -1, NONE = -2, BEFORE = -1;   // bucket: NONE = empty, BEFORE = &_M_before_begin
    int16_t next[MAXK];
    int16_t bucket[MAXB];
    int16_t head;
    int nb, count, next_resize;
    uint32_t nb_magic;   // ceil(2^32 / nb): mod(key) without a hardware divide (keys and nb < 2^16)

    PG2_DEV void set_nb(int n) { nb = n; nb_magic = (uint32_t)(0xffffffffu / (uint32_t)n) + 1u; }
    PG2_DEV int mod(int key) const {
        if (nb == 1) return 0;
        uint32_t q = (uint32_t)(((uint64_t)(uint32_t)key * nb_magic) >> 32);
        int r = key - (int)q * nb;
        return r >= nb ? r - nb : (r < 0 ? r + nb : r);
    }

    PG2_DEV_NOINLINE static int next_bkt(int n, int* next_resize) {
        const unsigned char fast[14] = { 2, 2, 2, 3, 5, 5, 7, 7, 11, 11, 11, 11, 13, 13 };
        if (n < 14) {
            if (n == 0) return 1;
            *next_resize = fast[n];
            return fast[n];
        }
        const short primes[] = { 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 103, 109, 113, 127, 137, 139,
                                 149, 157, 167, 179, 193, 199, 211, 227, 241, 257, 277, 293, 313, 337, 359, 383, 409, 439, 467, 503, 541,
                                 577, 619, 661, 709, 761, 823, 887, 953, 1031, 1109, 1193, 1289, 1381, 1493, 1613, 1741, 1879, 2029, 2179, 2357,
                                 2549, 2753, 2971, 3209, 3469, 3739, 4027, 4349, 4703, 5087 };
        int p = 5087;
        for (int i = 0; i < (int)(sizeof(primes) / sizeof(primes[0])); i++)
            if (primes[i] >= n) { p = primes[i]; break; }
        *next_resize = p;
        return p;
    }

    // Fresh set: persisted_nb = 1. After clear(): the bucket count seen last.
    PG2_DEV_NOINLINE void init(int persisted_nb) {
        set_nb(persisted_nb < 1 ? 1 : persisted_nb);
        next_resize = nb == 1 ? 0 : nb;
        count = 0;
        head = END;
        for (int i = 0; i < nb && i < MAXB; i++) bucket[i] = NONE;
    }

    PG2_DEV int16_t next_of(int16_t node) const { return node == BEFORE ? head : next[node]; }
    PG2_DEV void set_next(int16_t node, int16_t v) { if (node == BEFORE) head = v; else next[node] = v; }

    PG2_DEV static int mod_by(int key, int n, uint32_t magic) {
        if (n == 1) return 0;
        uint32_t q = (uint32_t)(((uint64_t)(uint32_t)key * magic) >> 32);
        int r = key - (int)q * n;
        return r >= n ? r - n : (r < 0 ? r + n : r);
    }

    PG2_DEV_NOINLINE void rehash(int newnb) {
        const uint32_t newmagic = (uint32_t)(0xffffffffu / (uint32_t)newnb) + 1u;
        for (int i = 0; i < newnb && i < MAXB; i++) bucket[i] = NONE;
        int16_t p = head;
        head = END;
        int bbegin_bkt = 0;
        while (p != END) {
            int16_t nxt = next[p];
            int bkt = mod_by(p, newnb, newmagic);
            if (bucket[bkt] == NONE) {
                next[p] = head;
                head = p;
                bucket[bkt] = BEFORE;
                if (next[p] != END) bucket[bbegin_bkt] = p;
                bbegin_bkt = bkt;
            } else {
                int16_t prev = bucket[bkt];
                next[p] = next_of(prev);
                set_next(prev, p);
            }
            p = nxt;
        }
        set_nb(newnb);
    }

    PG2_DEV_NOINLINE bool contains(int key) const {
        int16_t prev = bucket[mod(key)];
        if (prev == NONE) return false;
        for (int16_t p = next_of(prev); p != END && (mod(p)) == (mod(key)); p = next[p])
            if (p == key) return true;
        return false;
    }

    PG2_DEV_NOINLINE void insert(int key) {
        if (count > 0 && contains(key)) return;
        insert_new(key);
    }

    // insert a key the caller knows to be absent
    PG2_DEV_NOINLINE void insert_new(int key) {
        if (count + 1 > next_resize) {
            int min_bkts = (count + 1 > (next_resize ? 0 : 11)) ? count + 1 : (next_resize ? 0 : 11);   // load factor 1.0
            if (min_bkts >= nb) {
                int want = min_bkts + 1 > nb * 2 ? min_bkts + 1 : nb * 2;
                rehash(next_bkt(want, &next_resize));
            } else {
                next_resize = nb;
            }
        }
        int bkt = mod(key);
        if (bucket[bkt] != NONE) {
            int16_t prev = bucket[bkt];
            next[key] = next_of(prev);
            set_next(prev, (int16_t)key);
        } else {
            next[key] = head;
            head = (int16_t)key;
            if (next[key] != END) bucket[mod(next[key])] = (int16_t)key;
            bucket[bkt] = BEFORE;
        }
        count++;
    }

    PG2_DEV_NOINLINE void erase(int key) {
        if (count == 0) return;
        int bkt = mod(key);
        int16_t prev = bucket[bkt];
        if (prev == NONE) return;
        int16_t p = next_of(prev);
        while (p != END && p != key) {
            if ((mod(p)) != bkt) return;
            prev = p;
            p = next[p];
        }
        if (p == END) return;
        int16_t nxt = next[p];
        if (prev == bucket[bkt]) {
            // _M_remove_bucket_begin
            int next_bkt_i = nxt != END ? mod(nxt) : 0;
            if (nxt == END || next_bkt_i != bkt) {
                if (nxt != END) bucket[next_bkt_i] = bucket[bkt];
                bucket[bkt] = NONE;
            }
        } else if (nxt != END) {
            int next_bkt_i = mod(nxt);
            if (next_bkt_i != bkt) bucket[next_bkt_i] = prev;
        }
        set_next(prev, nxt);
        count--;
    }

    // iteration order -> out[0..count)
    template <class T>
    PG2_DEV_NOINLINE int order(T* out) const {
        int n = 0;
        for (int16_t p = head; p != END; p = next[p]) out[n++] = (T)p;
        return n;
    }
};

}  // namespace pg2
