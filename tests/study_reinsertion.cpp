// CPU study (planning aid, not a test and not part of the product): how much SAH does a reinsertion pass (Bittner et al. 2013 /
// Meister & Bittner 2018) take off the PLOC binary tree the GPU builder makes (csrc/bvh_build.cu step 3b)?
//   g++ -O2 -std=c++17 tests/study_reinsertion.cpp -o /tmp/study/reins && /tmp/study/reins boxes.bin [radius] [passes]
// boxes.bin: n x 6 float32 (lo xyz, hi xyz) per triangle.  Prints the summed inner-node half area / root half area (the node-visit
// term of the SAH) before and after each pass.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <queue>
#include <vector>

struct Box {
    float lo[3], hi[3];
    void grow(const Box& b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const { const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2]; return x * y + y * z + z * x; }
};
static Box merge(const Box& a, const Box& b) { Box r = a; r.grow(b); return r; }

static uint64_t spread(uint64_t x) {
    x &= 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

struct Tree {
    int n = 0;                       // leaves; nodes 0..2n-2, leaves are n-1+k
    std::vector<int> left, right, parent;
    std::vector<Box> box;
    int root = 0;
    bool leaf(int x) const { return x >= n - 1; }
    double inner_area() const {
        double s = 0;
        for (int i = 0; i < n - 1; i++) s += box[i].area();
        return s / box[root].area();
    }
    void refit_up(int x) {
        while (x >= 0) {
            box[x] = merge(box[left[x]], box[right[x]]);
            x = parent[x];
        }
    }
};

int main(int argc, char** argv) {
    if (argc < 2) return 1;
    const int radius = argc > 2 ? atoi(argv[2]) : 16, passes = argc > 3 ? atoi(argv[3]) : 3;
    FILE* f = fopen(argv[1], "rb");
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    const int n = (int)(bytes / 24);
    std::vector<Box> prim(n);
    if (fread(prim.data(), 24, n, f) != (size_t)n) return 1;
    fclose(f);
    Box scene = prim[0];
    for (auto& b : prim) scene.grow(b);
    std::vector<std::pair<uint64_t, int>> keys(n);
    for (int i = 0; i < n; i++) {
        uint64_t q[3];
        for (int a = 0; a < 3; a++) {
            const float c = 0.5f * (prim[i].lo[a] + prim[i].hi[a]);
            q[a] = (uint64_t)std::min(2097151.0, std::max(0.0, (double)(c - scene.lo[a]) / std::max(scene.hi[a] - scene.lo[a], 1e-30f) * 2097152.0));
        }
        keys[i] = {(spread(q[0]) << 2) | (spread(q[1]) << 1) | spread(q[2]), i};
    }
    std::sort(keys.begin(), keys.end());
    Tree t;
    t.n = n;
    t.left.assign(2 * n - 1, -1); t.right.assign(2 * n - 1, -1); t.parent.assign(2 * n - 1, -1); t.box.resize(2 * n - 1);
    // PLOC
    std::vector<int> cl(n), nxt;
    for (int i = 0; i < n; i++) { cl[i] = n - 1 + i; t.box[n - 1 + i] = prim[keys[i].second]; }
    int next_internal = n - 2;   // allocate from the back so that the root ends up at 0
    while (cl.size() > 1) {
        const int m = (int)cl.size();
        std::vector<int> nn(m, -1);
        for (int i = 0; i < m; i++) {
            float best = INFINITY;
            for (int j = std::max(0, i - radius); j <= std::min(m - 1, i + radius); j++) {
                if (j == i) continue;
                const float a = merge(t.box[cl[i]], t.box[cl[j]]).area();
                if (a < best) { best = a; nn[i] = j; }
            }
        }
        nxt.clear();
        for (int i = 0; i < m; i++) {
            const int j = nn[i];
            if (nn[j] == i) {
                if (i < j) {
                    const int id = next_internal--;
                    t.left[id] = cl[i]; t.right[id] = cl[j]; t.parent[cl[i]] = id; t.parent[cl[j]] = id;
                    t.box[id] = merge(t.box[cl[i]], t.box[cl[j]]);
                    nxt.push_back(id);
                }
            } else nxt.push_back(cl[i]);
        }
        cl.swap(nxt);
    }
    t.root = cl[0];
    printf("n = %d, PLOC radius %d: inner area / root area = %.4f\n", n, radius, t.inner_area());

    // reinsertion passes: nodes by decreasing area; remove x (parent p replaced by sibling s), find the position `best` minimising the
    // area increase along the insertion path (branch and bound over the whole tree), reinsert with p as the new common parent
    for (int pass = 0; pass < passes; pass++) {
        std::vector<int> order;
        for (int i = 0; i < 2 * n - 1; i++)
            if (i != t.root && t.parent[i] != t.root) order.push_back(i);
        std::sort(order.begin(), order.end(), [&](int a, int b) { return t.box[a].area() > t.box[b].area(); });
        long moved = 0;
        for (int x : order) {
            const int p = t.parent[x];
            if (p < 0 || p == t.root) continue;
            const int g = t.parent[p];
            const int s = t.left[p] == x ? t.right[p] : t.left[p];
            // remove
            (t.left[g] == p ? t.left[g] : t.right[g]) = s;
            t.parent[s] = g;
            t.refit_up(g);
            // search
            const Box bx = t.box[x];
            const float ax = bx.area();
            struct Item { float induced; int node; bool operator<(const Item& o) const { return induced > o.induced; } };
            std::priority_queue<Item> pq;
            pq.push({0.f, t.root});
            float best_cost = INFINITY;
            int best = -1;
            while (!pq.empty()) {
                const Item it = pq.top();
                pq.pop();
                if (it.induced + ax >= best_cost) break;
                const float direct = merge(t.box[it.node], bx).area();
                const float total = it.induced + direct;
                if (total < best_cost) { best_cost = total; best = it.node; }
                if (!t.leaf(it.node)) {
                    const float ind = total - t.box[it.node].area();
                    if (ind + ax < best_cost) { pq.push({ind, t.left[it.node]}); pq.push({ind, t.right[it.node]}); }
                }
            }
            // insert: p becomes the parent of {best, x}
            const int bp = t.parent[best];
            if (best != s || bp != g) moved++;
            t.left[p] = best; t.right[p] = x;
            t.parent[p] = bp;
            if (bp >= 0) (t.left[bp] == best ? t.left[bp] : t.right[bp]) = p;
            else t.root = p;
            t.parent[best] = p;
            t.parent[x] = p;
            t.refit_up(p);
        }
        printf("pass %d: %ld nodes moved, inner area / root area = %.4f\n", pass + 1, moved, t.inner_area());
    }
    return 0;
}
