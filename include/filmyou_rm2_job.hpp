// filmyou_rm2_job.hpp -- C++ host-side mirror of the reference's job interface over the C ABI
// (the reference is compiled Java; no JVM exists in this image, so the host side above the ABI is
// C++ here and ctypes in the tests).  Header only; link with -lfilmyou_rm2.
//
//   filmyou::RM2Job::run                      M/rm/RM2Job.java:76-100
//   filmyou::PreferenceSink::writePreference  M/rm/AbstractRM2Reducer.java:405-407
//   a failed job throws std::runtime_error    M/rm/RM2Job.java:265-268 ("... failed!")
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "filmyou_rm2.h"

namespace filmyou {

struct PreferenceSink {
    virtual ~PreferenceSink() = default;
    // same argument order and meaning as the reducer's abstract method
    virtual void writePreference(int userId, int itemId, double score, int cluster) = 0;
};

// what RM2HDFSReducer appends: (IntPairWritable(user,item), FloatWritable((float) score))
struct HDFSSink : PreferenceSink {
    struct Record { int32_t user, item; float score; };
    std::vector<Record> records;
    void writePreference(int userId, int itemId, double score, int) override {
        records.push_back(Record{userId, itemId, static_cast<float>(score)});
    }
};

struct RM2Conf {                      // RMRecommenderDriver.java:89-120
    double lambda = 0.1;
    int numberOfItems = 0;            // required
    int numberOfClusters = 0;         // required
    int numberOfRecommendations = 1000;
    int filterUsers = 0;
    int device = 0, shardRank = 0, shardCount = 1;
};

class RM2Job {
public:
    explicit RM2Job(const RM2Conf& conf) : conf_(conf) {
        fy_rm2_params p;
        fy_rm2_default_params(&p);
        p.lambda = conf.lambda; p.number_of_items = conf.numberOfItems; p.top_n = conf.numberOfRecommendations;
        p.filter_users = conf.filterUsers; p.device = conf.device; p.shard_rank = conf.shardRank; p.shard_count = conf.shardCount;
        const int rc = fy_rm2_create(&ctx_, &p);
        if (rc != FY_OK) throw std::runtime_error("fy_rm2_create failed: " + std::to_string(rc));
    }
    ~RM2Job() { fy_rm2_destroy(ctx_); }
    RM2Job(const RM2Job&) = delete;
    RM2Job& operator=(const RM2Job&) = delete;

    // ratings = the input SequenceFile's records; clustering / clusteringCount = the two cache files
    void run(const std::vector<int32_t>& user, const std::vector<int32_t>& item, const std::vector<float>& score,
             const std::vector<int32_t>& clusteringUser, const std::vector<int32_t>& clusteringCluster,
             std::vector<int32_t> clusteringCount, PreferenceSink& sink) {
        clusteringCount.resize(static_cast<size_t>(conf_.numberOfClusters), 0);
        check(fy_rm2_set_ratings(ctx_, user.data(), item.data(), score.data(), static_cast<int64_t>(user.size())), "ratings");
        check(fy_rm2_set_clustering(ctx_, clusteringUser.data(), clusteringCluster.data(),
                                    static_cast<int64_t>(clusteringUser.size()), clusteringCount.data(),
                                    conf_.numberOfClusters), "clustering");
        check(fy_rm2_run(ctx_), "RM2-3");
        userSum.assign(clusteringUser.size(), 0.0);
        itemColl.assign(static_cast<size_t>(fy_rm2_max_item(ctx_)) + 1, 0.0);
        check(fy_rm2_stats(ctx_, userSum.data(), itemColl.data(), &totalSum), "stats");
        const int64_t n = fy_rm2_result_count(ctx_);
        std::vector<int32_t> u(n), i(n), c(n);
        std::vector<double> s(n);
        check(fy_rm2_results(ctx_, u.data(), i.data(), s.data(), nullptr, c.data()), "results");
        for (int64_t k = 0; k < n; k++) sink.writePreference(u[k], i[k], s[k], c[k]);
    }

    std::vector<double> userSum, itemColl;   // rm2/userSum, rm2/itemColl
    double totalSum = 0.0;

private:
    void check(int rc, const char* job) {
        if (rc != FY_OK) throw std::runtime_error(std::string(job) + " failed! " + fy_rm2_last_error(ctx_));
    }
    RM2Conf conf_;
    fy_rm2_ctx* ctx_ = nullptr;
};

}  // namespace filmyou
