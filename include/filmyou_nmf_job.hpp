// filmyou_nmf_job.hpp -- C++ host-side mirror of the reference's clustering drivers over the C ABI of
// include/filmyou_nmf.h (header only; link with -lfilmyou_rm2).
//
//   filmyou::NMFDriver / PPCDriver::run          M/nmf/AbstractNMFDriver.java:92-141 (NMFDriver.java, ppc/PPCDriver.java)
//   filmyou::ClusterAssignment                   M/nmf/clustering/ClusterAssignmentJob.java:47-95 + CountClustersJob.java:41-80
//   a failed job throws std::runtime_error       ("... failed!", as the drivers do)
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "filmyou_nmf.h"

namespace filmyou {

struct NMFConf {                      // RMRecommenderDriver.java:89-120
    int numberOfUsers = 0, numberOfItems = 0, numberOfClusters = 0;   // required
    int numberOfIterations = 10;
    int normalizationFrequency = 12;
    int idBase = 1;                   // first user / item id (DataInitialization writes keys 1..rows)
    int device = 0;
};

struct ClusterAssignment {
    std::vector<int32_t> clustering;        // per user row: arg max of h_j   (the `clustering` file's values)
    std::vector<int32_t> clusteringCount;   // per cluster: users assigned    (the `clusteringCount` file)
};

class MatrixFactorizationDriver {
public:
    // H [numberOfUsers x k] and W [numberOfItems x k], row-major: the "H" / "W" options; empty = createInitialMatrices(seed)
    std::vector<double> H, W;

    MatrixFactorizationDriver(int mode, const NMFConf& conf) : conf_(conf) {
        fy_nmf_params p;
        fy_nmf_default_params(&p);
        p.mode = mode;
        p.number_of_users = conf.numberOfUsers; p.number_of_items = conf.numberOfItems; p.number_of_clusters = conf.numberOfClusters;
        p.number_of_iterations = conf.numberOfIterations; p.normalization_frequency = conf.normalizationFrequency;
        p.id_base = conf.idBase; p.device = conf.device;
        const int rc = fy_nmf_create(&ctx_, &p);
        if (rc != FY_OK) throw std::runtime_error("fy_nmf_create failed: " + std::to_string(rc));
    }
    ~MatrixFactorizationDriver() { fy_nmf_destroy(ctx_); }
    MatrixFactorizationDriver(const MatrixFactorizationDriver&) = delete;
    MatrixFactorizationDriver& operator=(const MatrixFactorizationDriver&) = delete;

    // ratings = the input SequenceFile's <(user, item), score> records; numberOfIterations x (H job, W job)
    void run(const std::vector<int32_t>& user, const std::vector<int32_t>& item, const std::vector<float>& score, uint64_t seed = 0) {
        check(fy_nmf_set_ratings(ctx_, user.data(), item.data(), score.data(), static_cast<int64_t>(user.size())), "ratings");
        if (!H.empty() || !W.empty()) {
            const size_t k = static_cast<size_t>(conf_.numberOfClusters);
            if (H.size() != static_cast<size_t>(conf_.numberOfUsers) * k || W.size() != static_cast<size_t>(conf_.numberOfItems) * k)
                throw std::runtime_error("H / W have the wrong shape");
            check(fy_nmf_set_factors(ctx_, H.data(), W.data()), "factors");
        } else {
            check(fy_nmf_init_random(ctx_, seed), "createInitialMatrices");
        }
        check(fy_nmf_run(ctx_), "matrix factorization");
        H.assign(static_cast<size_t>(conf_.numberOfUsers) * conf_.numberOfClusters, 0.0);
        W.assign(static_cast<size_t>(conf_.numberOfItems) * conf_.numberOfClusters, 0.0);
        check(fy_nmf_get_factors(ctx_, H.data(), W.data()), "factors");
    }

    ClusterAssignment assignClusters() {
        ClusterAssignment a;
        a.clustering.assign(static_cast<size_t>(conf_.numberOfUsers), 0);
        a.clusteringCount.assign(static_cast<size_t>(conf_.numberOfClusters), 0);
        check(fy_nmf_cluster_assignment(ctx_, a.clustering.data(), a.clusteringCount.data()), "ClusterAssignmentJob");
        return a;
    }

private:
    void check(int rc, const char* job) {
        if (rc != FY_OK) throw std::runtime_error(std::string(job) + " failed! " + fy_nmf_last_error(ctx_));
    }
    NMFConf conf_;
    fy_nmf_ctx* ctx_ = nullptr;
};

struct NMFDriver : MatrixFactorizationDriver { explicit NMFDriver(const NMFConf& c) : MatrixFactorizationDriver(0, c) {} };
struct PPCDriver : MatrixFactorizationDriver { explicit PPCDriver(const NMFConf& c) : MatrixFactorizationDriver(1, c) {} };

}  // namespace filmyou
