// frontier_pq.cpp -- TEST INFRASTRUCTURE ONLY (part of the CPU oracle).
//
// The frontier selection of the reference's RobotModel::GetExplorationFeatures
// (robot_model.h:383-401) is a std::priority_queue with FrontierCompare
// (typedefs.h:69-86): which of several equally valued candidates survive, and the
// order they are popped in, is decided by the standard library's heap algorithms.
// The oracle therefore runs the real std::priority_queue (this image's libstdc++,
// the library the reference links) rather than a restatement of it; the CUDA kernel's
// own heap (csrc/maps.cuh expl_*) is checked against this.
#include <queue>
#include <vector>

namespace {
struct Frontier {  // typedefs.h:69-74, the point replaced by the candidate's number
  int id;
  double value;
};
struct FrontierCompare {  // typedefs.h:79-83
  bool operator()(Frontier const &left, Frontier const &right) const { return left.value > right.value; }
};
typedef std::priority_queue<Frontier, std::vector<Frontier>, FrontierCompare> queue_t;
}  // namespace

// values[n] in the reference's candidate order; order_out[k] = candidate numbers as popped
// (worst first); returns how many were popped (min(n, num_frontiers)).
extern "C" int orc_frontier_select(const double *values, int n, int num_frontiers, int *order_out) {
  queue_t frontiers;
  for (int i = 0; i < n; ++i) {
    const double bcr = values[i];
    if (frontiers.size() < size_t(num_frontiers)) {  // robot_model.h:383-384
      frontiers.push(Frontier{i, bcr});
    } else {
      Frontier worst_frontier = frontiers.top();     // robot_model.h:386-390
      if (worst_frontier.value < bcr) {
        frontiers.pop();
        frontiers.push(Frontier{i, bcr});
      }
    }
  }
  int count = 0;
  while (!frontiers.empty()) {                       // robot_model.h:395-401
    order_out[count++] = frontiers.top().id;
    frontiers.pop();
  }
  return count;
}
