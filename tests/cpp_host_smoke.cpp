// Compiled and run by tests/test_gpu_parity.py::test_cpp_host_mirror: the C++ mirror of RM2Job over
// the C ABI on the 5x3 toy of T/testdata/RMTestData2.java (users 1..5, clusters {1,2}->0, {3,4,5}->1).
#include <cstdio>
#include "filmyou_rm2_job.hpp"

int main() {
    using namespace filmyou;
    // A[item][user] = { {5,0,0,1,2}, {4,3,1,0,0}, {0,0,2,4,5} }
    std::vector<int32_t> user = {1, 4, 5, 1, 2, 3, 3, 4, 5}, item = {1, 1, 1, 2, 2, 2, 3, 3, 3};
    std::vector<float> score = {5, 1, 2, 4, 3, 1, 2, 4, 5};
    RM2Conf conf; conf.lambda = 0.5; conf.numberOfItems = 3; conf.numberOfClusters = 2; conf.numberOfRecommendations = 10;
    RM2Job job(conf);
    HDFSSink sink;
    job.run(user, item, score, {1, 2, 3, 4, 5}, {0, 0, 1, 1, 1}, {2, 3}, sink);
    std::printf("totalSum %.1f\n", job.totalSum);
    for (double s : job.userSum) std::printf("userSum %.1f\n", s);
    for (const auto& r : sink.records) std::printf("rec %d %d %.6f\n", r.user, r.item, r.score);
    return 0;
}
