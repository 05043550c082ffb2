// train_state.cpp -- see train_state.hpp
#include "train_state.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>

namespace spchost {

static bool write_tree(const std::string& path, const std::vector<spc_tree_node>& t) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return false;
    for (const spc_tree_node& n : t) {
        if (n.leaf) {
            fprintf(f, "1 %d\n", n.label);
        } else {
            fprintf(f, "0 %d %d %.9g %.9g %.9g", n.label, n.type, n.mid.x, n.mid.y, n.mid.z);
            for (int c = 0; c < 8; c++) fprintf(f, " %d", n.child[c]);
            fputc('\n', f);
        }
    }
    return fclose(f) == 0;
}

static bool read_tree(const std::string& path, std::vector<spc_tree_node>& t) {
    std::ifstream in(path.c_str());
    if (!in) return false;
    t.clear();
    int leaf;
    while (in >> leaf) {
        spc_tree_node n;
        memset(&n, 0, sizeof(n));
        in >> n.label;
        n.leaf = (uint8_t)(leaf != 0);
        if (!leaf) {
            in >> n.type >> n.mid.x >> n.mid.y >> n.mid.z;
            for (int c = 0; c < 8; c++) in >> n.child[c];
        }
        if (!in) return false;
        t.push_back(n);
    }
    return !t.empty();
}

static bool write_floats(const std::string& path, const std::vector<float>& v, size_t per_line) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return false;
    for (size_t i = 0; i < v.size(); i++) fprintf(f, "%.9g%c", v[i], (i + 1) % per_line == 0 ? '\n' : ' ');
    return fclose(f) == 0;
}

static bool read_floats(const std::string& path, std::vector<float>& v) {
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    v.clear();
    float x;
    while (fscanf(f, "%f", &x) == 1) v.push_back(x);
    fclose(f);
    return true;
}

bool save_train_state(const std::string& p, const TrainState& s) {
    return write_tree(p + "tree_eye.txt", s.eye_tree) && write_tree(p + "tree_light.txt", s.light_tree) && write_floats(p + "Q.txt", s.Q, 1) &&
           write_floats(p + "E.txt", s.gamma, s.Q.empty() ? 1 : s.Q.size());
}

bool load_train_state(const std::string& p, int K, TrainState& s, std::string& err) {
    if (!read_tree(p + "tree_eye.txt", s.eye_tree) || !read_tree(p + "tree_light.txt", s.light_tree)) {
        err = "cannot read " + p + "tree_eye.txt / tree_light.txt";
        return false;
    }
    if (!read_floats(p + "Q.txt", s.Q) || (int)s.Q.size() != K) {
        err = "cannot read " + p + "Q.txt (expected " + std::to_string(K) + " values)";
        return false;
    }
    if (!read_floats(p + "E.txt", s.gamma) || s.gamma.size() != (size_t)K * K) {
        err = "cannot read " + p + "E.txt (expected K*K values)";
        return false;
    }
    return true;
}

}  // namespace spchost
