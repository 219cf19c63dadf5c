// CPU harness of the FP32 pose fast path (csrc/pose_fast.cuh is __host__ __device__): test infrastructure.
// Built by tests/test_pose_fast_host.py with g++ and compared against oracle/pose_oracle.py.
#include <cmath>
#include <cstdint>
#include "../humanpose-and-urbanscene-matching_b200/csrc/pose_fast.cuh"

// precise != 0: the same formulas in double with zero-width bands (the yardstick); `certain` then only says that the
// pose did not need one of the reference's corner-case paths.
extern "C" int pose_fast_host(const float* query, const float* db, int64_t n, int query_is_model, const double* thresholds,
                              int precise, uint8_t* match, double* score, uint8_t* certain) {
  hpusm::PoseQueryPrep pq;
  hpusm::pose_query_prep(query, pq);
  hpusm::PoseFastParams prm;
  prm.d_torso = thresholds[0]; prm.tan_torso = std::tan(thresholds[1] / 57.3);
  prm.d_legs = thresholds[2]; prm.tan_legs = std::tan(thresholds[3] / 57.3);
  prm.d_shoulder = thresholds[4];
  for (int64_t i = 0; i < n; ++i) {
    float px[18], py[18];
    for (int j = 0; j < 18; ++j) { px[j] = db[i * 36 + 2 * j]; py[j] = db[i * 36 + 2 * j + 1]; }
    uint8_t m = 0;
    double s = NAN;
    bool c = false;
    if (pq.usable && precise) {
      c = query_is_model ? hpusm::pose_fast_decide<double, true>(pq, px, py, prm, 0.0, 0.0, m, s)
                         : hpusm::pose_fast_decide<double, false>(pq, px, py, prm, 0.0, 0.0, m, s);
    } else if (pq.usable) {
      float sf = NAN;
      c = query_is_model ? hpusm::pose_fast_decide<float, true>(pq, px, py, prm, hpusm::PF_DIST_BAND, hpusm::PF_ROT_BAND, m, sf)
                         : hpusm::pose_fast_decide<float, false>(pq, px, py, prm, hpusm::PF_DIST_BAND, hpusm::PF_ROT_BAND, m, sf);
      s = sf;
    }
    match[i] = m; score[i] = s; certain[i] = c ? 1 : 0;
  }
  return pq.usable;
}
