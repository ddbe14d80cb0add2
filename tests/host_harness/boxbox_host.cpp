// TEST HARNESS (tests/): runs the product's box-box geometry (genesis_b200/csrc/gsr_boxbox.cuh, compiled for the host) next to the
// oracle's restatement (oracle/box_box.hpp, fp32) on random box pairs and prints the disagreements.
//   usage: boxbox_host <n_pairs> <seed>   -> "pairs N colliding M mismatches K maxerr X"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#define __noinline__
#endif
#include "../../genesis_b200/csrc/gsr_boxbox.cuh"
#include "../../oracle/box_box.hpp"

struct Out {
  int n = 0;
  f3 pos[5], nrm[5];
  float pen[5];
  void add(f3 normal, f3 p, float pe) { if (n < 5) { nrm[n] = normal; pos[n] = p; pen[n] = pe; n++; } }
};

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 1000;
  std::mt19937 rng(argc > 2 ? atoi(argv[2]) : 0);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  const float EPS = 1.1920929e-07f;
  int colliding = 0, mismatches = 0;
  double maxerr = 0;
  for (int it = 0; it < N; it++) {
    f3 s1(0.05f + 0.1f * fabsf(U(rng)), 0.05f + 0.1f * fabsf(U(rng)), 0.05f + 0.1f * fabsf(U(rng)));
    f3 s2(0.05f + 0.1f * fabsf(U(rng)), 0.05f + 0.1f * fabsf(U(rng)), 0.05f + 0.1f * fabsf(U(rng)));
    f4 q1(U(rng), U(rng), U(rng), U(rng)), q2(U(rng), U(rng), U(rng), U(rng));
    if (it % 4 == 0) { q1 = f4(1, 0, 0, 0); }                        // axis-aligned lower box
    if (it % 8 == 0) { q2 = f4(cosf(0.3f * it), 0, 0, sinf(0.3f * it)); }  // yaw only: face-face with parallel edges
    float n1 = sqrtf(q1.w * q1.w + q1.x * q1.x + q1.y * q1.y + q1.z * q1.z), n2 = sqrtf(q2.w * q2.w + q2.x * q2.x + q2.y * q2.y + q2.z * q2.z);
    q1 = f4(q1.w / n1, q1.x / n1, q1.y / n1, q1.z / n1); q2 = f4(q2.w / n2, q2.x / n2, q2.y / n2, q2.z / n2);
    f3 p1(0.1f * U(rng), 0.1f * U(rng), 0.1f * U(rng)), p2(0.14f * U(rng), 0.14f * U(rng), 0.14f * U(rng));
    Out out;
    bb_contacts(p1, quat_to_R(q1, EPS), s1, p2, quat_to_R(q2, EPS), s2, EPS, 5, out);
    orc::BoxBoxContacts<float> bf;
    const bool cf = orc::box_box_detect<float>(orc::V3<float>(p1.x, p1.y, p1.z), orc::Q4<float>(q1.w, q1.x, q1.y, q1.z), orc::V3<float>(s1.x, s1.y, s1.z),
                                               orc::V3<float>(p2.x, p2.y, p2.z), orc::Q4<float>(q2.w, q2.x, q2.y, q2.z), orc::V3<float>(s2.x, s2.y, s2.z), EPS, bf);
    if (!cf) bf.n = 0;
    // the oracle's duplicate filter and cap (rigid_oracle.hpp box_box_contact)
    int kept = 0, idx[5];
    for (int i = 0; i < bf.n && kept < 5; i++) {
      bool valid = true;
      for (int j = 0; j < kept; j++) {
        auto d = bf.pos[i] - bf.pos[idx[j]];
        if (fabsf(d.x) < EPS && fabsf(d.y) < EPS && fabsf(d.z) < EPS) valid = false;
      }
      if (valid) idx[kept++] = i;
    }
    if (cf) { colliding++; }
    bool bad = kept != out.n;
    double err = 0;
    for (int k = 0; k < kept && !bad; k++) {
      const int i = idx[k];
      err = fmax(err, fabs(out.pos[k].x - bf.pos[i].x)); err = fmax(err, fabs(out.pos[k].y - bf.pos[i].y)); err = fmax(err, fabs(out.pos[k].z - bf.pos[i].z));
      err = fmax(err, fabs(out.nrm[k].x + bf.normal.x)); err = fmax(err, fabs(out.nrm[k].y + bf.normal.y)); err = fmax(err, fabs(out.nrm[k].z + bf.normal.z));
      err = fmax(err, fabs(out.pen[k] + bf.dist[i]));
    }
    if (err > 2e-5) bad = true;
    maxerr = fmax(maxerr, bad ? 0 : err);
    if (bad) { mismatches++; if (mismatches <= 5) fprintf(stderr, "pair %d: product n=%d oracle n=%d (candidates %d) err %.3e\n", it, out.n, kept, bf.n, err); }
  }
  printf("pairs %d colliding %d mismatches %d maxerr %.3e\n", N, colliding, mismatches, maxerr);
  return mismatches ? 1 : 0;
}
