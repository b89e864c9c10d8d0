// include/Timings.h -- per-stage timings of one multiply, same fields as the reference's
// include/Timings.h:4-18 (milliseconds).  Stages that no longer exist in the B200 pipeline
// (load balancing, global hash maps, separate sorting, cleanup) stay in the struct and read 0.
#pragma once

struct Timings {
    bool measureAll = false;
    bool measureCompleteTime = false;
    float init = 0.0f;
    float countProducts = 0.0f;
    float loadBalanceCounting = 0.0f;
    float globalMapsCounting = 0.0f;
    float spGEMMCounting = 0.0f;
    float allocC = 0.0f;
    float loadBalanceNumeric = 0.0f;
    float globalMapsNumeric = 0.0f;
    float spGEMMNumeric = 0.0f;
    float sorting = 0.0f;
    float cleanup = 0.0f;
    float complete = 0.0f;

    static constexpr int kStages = 12;
    float *stage(int i) { return &init + i; }
    const float *stage(int i) const { return &init + i; }

    void operator+=(const Timings &o)
    {
        for (int i = 0; i < kStages; ++i) *stage(i) += *o.stage(i);
    }
    void operator/=(const float &d)
    {
        for (int i = 0; i < kStages; ++i) *stage(i) /= d;
    }
};
static_assert(sizeof(Timings) == 2 * sizeof(bool) + 2 /* padding */ + Timings::kStages * sizeof(float),
              "the stage fields must be contiguous floats");
