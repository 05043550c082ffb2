// train_state.hpp -- the trained subspace state on disk: the text files the reference reads back for debugging,
// classTree::tree_load ("tree_eye.txt" / "tree_light.txt", decisionTree/classTree_host.h:15-60), MyThrustOp::load_Q_file
// ("Q.txt") and load_Gamma_file ("E.txt") (cuda_thrust/device_thrust.cu:3347-3404).  Same token order, so files written here
// load in the reference and vice versa; floats are written with 9 significant digits (binary32 round trip).
//   <prefix>tree_eye.txt / <prefix>tree_light.txt : per node "leaf label" and, for inner nodes, "type mid.x mid.y mid.z child[0..7]"
//   <prefix>Q.txt : K floats;   <prefix>E.txt : K*K floats, row = eye subspace
#pragma once
#include <string>
#include <vector>

#include "spcbpt_b200.h"

namespace spchost {

struct TrainState {
    std::vector<spc_tree_node> eye_tree, light_tree;
    std::vector<float> Q, gamma;   // K and K*K
};

bool save_train_state(const std::string& prefix, const TrainState& s);
bool load_train_state(const std::string& prefix, int K, TrainState& s, std::string& err);

}  // namespace spchost
