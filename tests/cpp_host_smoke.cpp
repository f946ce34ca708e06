// Compiled and run by tests/test_gpu_parity.py::test_cpp_host_mirror: the C++ mirror of RM2Job over
// the C ABI on the 5x3 toy of T/testdata/RMTestData2.java (users 1..5, clusters {1,2}->0, {3,4,5}->1).
#include <cstdio>
#include "filmyou_nmf_job.hpp"
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

    // the clustering step: one PPC iteration on the same 3 x 5 toy (items x users), k = 2, then arg-max assignment
    NMFConf nconf; nconf.numberOfUsers = 5; nconf.numberOfItems = 3; nconf.numberOfClusters = 2; nconf.numberOfIterations = 1;
    PPCDriver ppc(nconf);
    ppc.H = {0.2, 0.8, 0.6, 0.4, 0.5, 0.5, 0.9, 0.1, 0.3, 0.7};
    ppc.W = {0.7, 0.3, 0.4, 0.6, 0.1, 0.9};
    ppc.run(user, item, score);
    for (double h : ppc.H) std::printf("H %.17g\n", h);
    const ClusterAssignment a = ppc.assignClusters();
    for (int32_t c : a.clustering) std::printf("cluster %d\n", c);
    std::printf("count %d %d\n", a.clusteringCount[0], a.clusteringCount[1]);
    return 0;
}
