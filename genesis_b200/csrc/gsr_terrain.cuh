// gsr_terrain.cuh -- a convex geom against the height-field terrain: MPR against the triangular prisms of the cells under the
// geom's bounding box in the terrain's frame (genesis/engine/solvers/rigid/collider/narrowphase.py:346-517 func_contact_mpr_terrain,
// support_field.py:268-285 _func_support_prism, collider.py:541-561 terrain tables).  One lane per pair; the prism (6 points) and
// the portal live in the lane's registers / local memory.  Compiled into the k_step<G, true, true> variants only.
#pragma once

// the pair as MPR sees it: geom A posed in the terrain's frame, B = the current prism (bottom triangle [0..3), top [3..6))
template <int G>
struct TerrainSupport {
  const Ctx<G>& C;
  int ga;
  f3 pa;
  f4 qa;
  const f3* prism;
  __device__ __forceinline__ void operator()(f3 d, f3& v, f3& v1, f3& v2) const {
    v1 = dev_support(C, d, ga, pa, qa);
    const f3 nd = -d;
    const int i0 = nd.z < 0.f ? 0 : 3;
    int ib = i0;
    float best = dot(prism[i0], nd);
    for (int i = i0 + 1; i < i0 + 3; i++) { const float t = dot(prism[i], nd); if (t > best) { ib = i; best = t; } }
    v2 = prism[ib];
    v = v1 - v2;
  }
};

template <int G>
__device__ __noinline__ void dev_terrain(const Ctx<G>& C, LaneContacts& out, int ga, int gb) {
  f3 ga_pos, gb_pos;
  f4 ga_quat, gb_quat;
  geom_pose(C, ga, ga_pos, ga_quat);
  geom_pose(C, gb, gb_pos, gb_quat);
  const float tolerance = dev_tolerance(C, ga, gb);
  const f4 inv_b = quat_inv(gb_quat);
  const f3 pa = rotate(ga_pos - gb_pos, inv_b);
  const f4 qa = quat_compose(ga_quat, inv_b);
  const f3 center_a = transform(f3(GM(g_center[ga][0]), GM(g_center[ga][1]), GM(g_center[ga][2])), pa, qa);
  // bounding box of A along the terrain's axes, from six support queries (NP:380-397)
  float hi[3], lo[3];
  for (int ax = 0; ax < 3; ax++) {
    f3 d;
    set_comp(d, ax, 1.f);
    hi[ax] = comp(dev_support(C, d, ga, pa, qa), ax);
    set_comp(d, ax, -1.f);
    lo[ax] = comp(dev_support(C, d, ga, pa, qa), ax);
  }
  for (int i = 0; i < 3; i++)
    if (CM(terrain_maxmin[i]) < lo[i] || CM(terrain_maxmin[i + 3]) > hi[i]) return;
  const float sh = CM(terrain_hs), x0 = CM(terrain_maxmin[3]), y0 = CM(terrain_maxmin[4]);
  int r_min = (int)floorf((lo[0] - x0) / sh), r_max = (int)ceilf((hi[0] - x0) / sh);
  int c_min = (int)floorf((lo[1] - y0) / sh), c_max = (int)ceilf((hi[1] - y0) / sh);
  r_min = max(0, r_min); c_min = max(0, c_min);
  r_max = min(CM(terrain_rows) - 1, r_max); c_max = min(CM(terrain_cols) - 1, c_max);
  const int ncols = CM(terrain_cols), ncpp = min(c_m.n_contacts_per_pair, 5);
  const float* hf = CM(terrain_hf);
  f3 prism[6];
  for (int i = 0; i < 3; i++) prism[i].z = CM(terrain_maxmin[5]);
  const TerrainSupport<G> sup{C, ga, pa, qa, prism};
  for (int r = r_min; r < r_max; r++) {
    int nvert = 0;
    for (int c = c_min; c <= c_max; c++)
      for (int i = 0; i < 2; i++) {
        if (out.n >= ncpp) continue;
        nvert++;
        {  // NP:499-517 func_add_prism_vert: the strip advances by one vertex
          const float x = sh * (float)(r + i) + x0, y = sh * (float)c + y0, z = __ldg(hf + (size_t)(r + i) * ncols + c);
          prism[0] = prism[1]; prism[1] = prism[2]; prism[3] = prism[4]; prism[4] = prism[5];
          prism[2].x = x; prism[5].x = x; prism[2].y = y; prism[5].y = y; prism[5].z = z;
        }
        if (nvert > 2 && (prism[3].z >= lo[2] || prism[4].z >= lo[2] || prism[5].z >= lo[2])) {
          f3 center_b;
          for (int k = 0; k < 6; k++) center_b += prism[k];
          center_b = center_b / 6.0f;
          bool is_col, tetra;
          f3 normal, cpos;
          float pen;
          Simplex sx;
          dev_mpr_from_centers(C, sup, center_a, center_b, is_col, normal, pen, cpos, sx, tetra);
          if (is_col) {
            normal = rotate(normal, gb_quat);
            cpos = rotate(cpos, gb_quat) + gb_pos;
            bool valid = true;
            for (int j = 0; j < out.n; j++)
              if (norm(cpos - out.pos[j]) < tolerance) valid = false;
            if (valid) out.add(normal, cpos, pen);
          }
        }
      }
  }
}
