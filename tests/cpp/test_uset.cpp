// TEST: pg2::USet iteration order == std::unordered_set<int> (libstdc++) under random
// insert / erase / clear / copy-free workloads, and the std::sort permutation table.
#define PG2_HOSTSIM 1
#include <stdio.h>
#include <algorithm>
#include <random>
#include <unordered_set>
#include <vector>
#include "../../procgen2_b200/csrc/pg2_uset.cuh"

using namespace pg2;

int main() {
    std::mt19937 rng(123);
    int checks = 0;
    for (int trial = 0; trial < 400; trial++) {
        std::unordered_set<int> ref;
        static USet<1000, 1200> us;
        us.init(1);
        int episodes = 1 + rng() % 6;
        for (int ep = 0; ep < episodes; ep++) {
            int ops = rng() % 300;
            int keyspace = 1 + rng() % (trial % 2 ? 200 : 1000);
            bool ascending = rng() % 2;
            int nextkey = 0;
            for (int o = 0; o < ops; o++) {
                int r = rng() % 10;
                if (r < 7) {
                    int k = ascending ? (nextkey++ % 1000) : (int)(rng() % keyspace);
                    ref.insert(k); us.insert(k);
                } else {
                    int k = rng() % keyspace;
                    ref.erase(k); us.erase(k);
                }
                if (o % 7 == 0 || o == ops - 1) {
                    std::vector<int> a(ref.begin(), ref.end());
                    std::vector<int> b(1000);
                    int n = us.order(b.data());
                    b.resize(n);
                    checks++;
                    if (a != b || (int)ref.bucket_count() != us.nb) {
                        printf("MISMATCH trial %d ep %d op %d: sizes %zu %d buckets %zu %d\n", trial, ep, o, a.size(), n, ref.bucket_count(), us.nb);
                        return 1;
                    }
                }
            }
            ref.clear();
            us.init(us.nb);   // clear(): keeps the bucket count
        }
    }
    // large generator-local sets (caveflyer / jumper best_room: up to 1600 keys, fresh set, insert only)
    for (int trial = 0; trial < 40; trial++) {
        std::unordered_set<int> ref;
        static USet<1600, 2400> big;
        big.init(1);
        int n = 200 + rng() % 1400;
        std::vector<int> keys(1600);
        for (int i = 0; i < 1600; i++) keys[i] = i;
        std::shuffle(keys.begin(), keys.end(), rng);
        for (int i = 0; i < n; i++) { ref.insert(keys[i]); big.insert(keys[i]); }
        std::vector<int> a(ref.begin(), ref.end());
        std::vector<int> b(1600);
        int m = big.order(b.data());
        b.resize(m);
        checks++;
        if (a != b || (int)ref.bucket_count() != big.nb) {
            printf("MISMATCH big trial %d: sizes %zu %d buckets %zu %d\n", trial, a.size(), m, ref.bucket_count(), big.nb);
            return 1;
        }
    }
    printf("OK %d checks\n", checks);
    return 0;
}
