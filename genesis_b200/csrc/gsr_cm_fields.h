// GENERATED (tools/specialize.py --fields): CM(x) resolves to CMX_<first token of x>; by default the field of the __constant__ model,
// in a model-specialised build (GSR_SPEC_HEADER) the scalars that are fixed at build time are literals.
#pragma once
#ifndef CMX_batch_dofs
#define CMX_batch_dofs (c_m.batch_dofs)
#endif
#ifndef CMX_box_box_detection
#define CMX_box_box_detection (c_m.box_box_detection)
#endif
#ifndef CMX_ccd_eps
#define CMX_ccd_eps (c_m.ccd_eps)
#endif
#ifndef CMX_ccd_iterations
#define CMX_ccd_iterations (c_m.ccd_iterations)
#endif
#ifndef CMX_ccd_tolerance
#define CMX_ccd_tolerance (c_m.ccd_tolerance)
#endif
#ifndef CMX_d_act_bias
#define CMX_d_act_bias (c_m.d_act_bias)
#endif
#ifndef CMX_d_armature
#define CMX_d_armature (c_m.d_armature)
#endif
#ifndef CMX_d_damping
#define CMX_d_damping (c_m.d_damping)
#endif
#ifndef CMX_d_entity
#define CMX_d_entity (c_m.d_entity)
#endif
#ifndef CMX_d_force_range
#define CMX_d_force_range (c_m.d_force_range)
#endif
#ifndef CMX_d_frictionloss
#define CMX_d_frictionloss (c_m.d_frictionloss)
#endif
#ifndef CMX_d_invweight
#define CMX_d_invweight (c_m.d_invweight)
#endif
#ifndef CMX_d_joint
#define CMX_d_joint (c_m.d_joint)
#endif
#ifndef CMX_d_kp
#define CMX_d_kp (c_m.d_kp)
#endif
#ifndef CMX_d_limit
#define CMX_d_limit (c_m.d_limit)
#endif
#ifndef CMX_d_link
#define CMX_d_link (c_m.d_link)
#endif
#ifndef CMX_d_motion_ang
#define CMX_d_motion_ang (c_m.d_motion_ang)
#endif
#ifndef CMX_d_motion_vel
#define CMX_d_motion_vel (c_m.d_motion_vel)
#endif
#ifndef CMX_d_stiffness
#define CMX_d_stiffness (c_m.d_stiffness)
#endif
#ifndef CMX_diagnostics
#define CMX_diagnostics (c_m.diagnostics)
#endif
#ifndef CMX_disable_constraint
#define CMX_disable_constraint (c_m.disable_constraint)
#endif
#ifndef CMX_dt
#define CMX_dt (c_m.dt)
#endif
#ifndef CMX_e_dof_end
#define CMX_e_dof_end (c_m.e_dof_end)
#endif
#ifndef CMX_e_dof_start
#define CMX_e_dof_start (c_m.e_dof_start)
#endif
#ifndef CMX_e_geom_end
#define CMX_e_geom_end (c_m.e_geom_end)
#endif
#ifndef CMX_e_geom_start
#define CMX_e_geom_start (c_m.e_geom_start)
#endif
#ifndef CMX_e_gravcomp
#define CMX_e_gravcomp (c_m.e_gravcomp)
#endif
#ifndef CMX_e_link_end
#define CMX_e_link_end (c_m.e_link_end)
#endif
#ifndef CMX_e_link_start
#define CMX_e_link_start (c_m.e_link_start)
#endif
#ifndef CMX_enable_collision
#define CMX_enable_collision (c_m.enable_collision)
#endif
#ifndef CMX_enable_joint_limit
#define CMX_enable_joint_limit (c_m.enable_joint_limit)
#endif
#ifndef CMX_enable_multi_contact
#define CMX_enable_multi_contact (c_m.enable_multi_contact)
#endif
#ifndef CMX_eps
#define CMX_eps (c_m.eps)
#endif
#ifndef CMX_eq_data
#define CMX_eq_data (c_m.eq_data)
#endif
#ifndef CMX_eq_obj1
#define CMX_eq_obj1 (c_m.eq_obj1)
#endif
#ifndef CMX_eq_obj2
#define CMX_eq_obj2 (c_m.eq_obj2)
#endif
#ifndef CMX_eq_sol
#define CMX_eq_sol (c_m.eq_sol)
#endif
#ifndef CMX_eq_type
#define CMX_eq_type (c_m.eq_type)
#endif
#ifndef CMX_fac_start
#define CMX_fac_start (c_m.fac_start)
#endif
#ifndef CMX_fac_tasks
#define CMX_fac_tasks (c_m.fac_tasks)
#endif
#ifndef CMX_g_aabb_hi
#define CMX_g_aabb_hi (c_m.g_aabb_hi)
#endif
#ifndef CMX_g_aabb_lo
#define CMX_g_aabb_lo (c_m.g_aabb_lo)
#endif
#ifndef CMX_g_center
#define CMX_g_center (c_m.g_center)
#endif
#ifndef CMX_g_data
#define CMX_g_data (c_m.g_data)
#endif
#ifndef CMX_g_friction
#define CMX_g_friction (c_m.g_friction)
#endif
#ifndef CMX_g_is_fixed
#define CMX_g_is_fixed (c_m.g_is_fixed)
#endif
#ifndef CMX_g_link
#define CMX_g_link (c_m.g_link)
#endif
#ifndef CMX_g_pos
#define CMX_g_pos (c_m.g_pos)
#endif
#ifndef CMX_g_quat
#define CMX_g_quat (c_m.g_quat)
#endif
#ifndef CMX_g_sol
#define CMX_g_sol (c_m.g_sol)
#endif
#ifndef CMX_g_support_start
#define CMX_g_support_start (c_m.g_support_start)
#endif
#ifndef CMX_g_type
#define CMX_g_type (c_m.g_type)
#endif
#ifndef CMX_g_vert_end
#define CMX_g_vert_end (c_m.g_vert_end)
#endif
#ifndef CMX_g_vert_start
#define CMX_g_vert_start (c_m.g_vert_start)
#endif
#ifndef CMX_gjk_fallback
#define CMX_gjk_fallback (c_m.gjk_fallback)
#endif
#ifndef CMX_gravity
#define CMX_gravity (c_m.gravity)
#endif
#ifndef CMX_has_box_box
#define CMX_has_box_box (c_m.has_box_box)
#endif
#ifndef CMX_has_cc
#define CMX_has_cc (c_m.has_cc)
#endif
#ifndef CMX_has_mpr
#define CMX_has_mpr (c_m.has_mpr)
#endif
#ifndef CMX_integrator
#define CMX_integrator (c_m.integrator)
#endif
#ifndef CMX_iterations
#define CMX_iterations (c_m.iterations)
#endif
#ifndef CMX_j_dof_start
#define CMX_j_dof_start (c_m.j_dof_start)
#endif
#ifndef CMX_j_n_dofs
#define CMX_j_n_dofs (c_m.j_n_dofs)
#endif
#ifndef CMX_j_pos
#define CMX_j_pos (c_m.j_pos)
#endif
#ifndef CMX_j_q_start
#define CMX_j_q_start (c_m.j_q_start)
#endif
#ifndef CMX_j_sol
#define CMX_j_sol (c_m.j_sol)
#endif
#ifndef CMX_j_type
#define CMX_j_type (c_m.j_type)
#endif
#ifndef CMX_l_ancmask
#define CMX_l_ancmask (c_m.l_ancmask)
#endif
#ifndef CMX_l_depth
#define CMX_l_depth (c_m.l_depth)
#endif
#ifndef CMX_l_descmask
#define CMX_l_descmask (c_m.l_descmask)
#endif
#ifndef CMX_l_dof_end
#define CMX_l_dof_end (c_m.l_dof_end)
#endif
#ifndef CMX_l_dof_start
#define CMX_l_dof_start (c_m.l_dof_start)
#endif
#ifndef CMX_l_dofmask
#define CMX_l_dofmask (c_m.l_dofmask)
#endif
#ifndef CMX_l_entity
#define CMX_l_entity (c_m.l_entity)
#endif
#ifndef CMX_l_inertia
#define CMX_l_inertia (c_m.l_inertia)
#endif
#ifndef CMX_l_invweight
#define CMX_l_invweight (c_m.l_invweight)
#endif
#ifndef CMX_l_invweight_rot
#define CMX_l_invweight_rot (c_m.l_invweight_rot)
#endif
#ifndef CMX_l_ipos
#define CMX_l_ipos (c_m.l_ipos)
#endif
#ifndef CMX_l_iquat
#define CMX_l_iquat (c_m.l_iquat)
#endif
#ifndef CMX_l_is_fixed
#define CMX_l_is_fixed (c_m.l_is_fixed)
#endif
#ifndef CMX_l_joint_end
#define CMX_l_joint_end (c_m.l_joint_end)
#endif
#ifndef CMX_l_joint_start
#define CMX_l_joint_start (c_m.l_joint_start)
#endif
#ifndef CMX_l_mass
#define CMX_l_mass (c_m.l_mass)
#endif
#ifndef CMX_l_parent
#define CMX_l_parent (c_m.l_parent)
#endif
#ifndef CMX_l_pos
#define CMX_l_pos (c_m.l_pos)
#endif
#ifndef CMX_l_q_end
#define CMX_l_q_end (c_m.l_q_end)
#endif
#ifndef CMX_l_q_start
#define CMX_l_q_start (c_m.l_q_start)
#endif
#ifndef CMX_l_quat
#define CMX_l_quat (c_m.l_quat)
#endif
#ifndef CMX_l_root
#define CMX_l_root (c_m.l_root)
#endif
#ifndef CMX_l_root_mass
#define CMX_l_root_mass (c_m.l_root_mass)
#endif
#ifndef CMX_l_rootslot
#define CMX_l_rootslot (c_m.l_rootslot)
#endif
#ifndef CMX_l_submass
#define CMX_l_submass (c_m.l_submass)
#endif
#ifndef CMX_len_constraints
#define CMX_len_constraints (c_m.len_constraints)
#endif
#ifndef CMX_ls_iterations
#define CMX_ls_iterations (c_m.ls_iterations)
#endif
#ifndef CMX_ls_tolerance
#define CMX_ls_tolerance (c_m.ls_tolerance)
#endif
#ifndef CMX_max_broad_pairs
#define CMX_max_broad_pairs (c_m.max_broad_pairs)
#endif
#ifndef CMX_max_contact_pairs
#define CMX_max_contact_pairs (c_m.max_contact_pairs)
#endif
#ifndef CMX_max_depth
#define CMX_max_depth (c_m.max_depth)
#endif
#ifndef CMX_mc_perturbation
#define CMX_mc_perturbation (c_m.mc_perturbation)
#endif
#ifndef CMX_mc_tolerance
#define CMX_mc_tolerance (c_m.mc_tolerance)
#endif
#ifndef CMX_meaninertia
#define CMX_meaninertia (c_m.meaninertia)
#endif
#ifndef CMX_nD
#define CMX_nD (c_m.nD)
#endif
#ifndef CMX_nE
#define CMX_nE (c_m.nE)
#endif
#ifndef CMX_nEq
#define CMX_nEq (c_m.nEq)
#endif
#ifndef CMX_nG
#define CMX_nG (c_m.nG)
#endif
#ifndef CMX_nJ
#define CMX_nJ (c_m.nJ)
#endif
#ifndef CMX_nL
#define CMX_nL (c_m.nL)
#endif
#ifndef CMX_nP
#define CMX_nP (c_m.nP)
#endif
#ifndef CMX_nQ
#define CMX_nQ (c_m.nQ)
#endif
#ifndef CMX_nV
#define CMX_nV (c_m.nV)
#endif
#ifndef CMX_n_cache_words
#define CMX_n_cache_words (c_m.n_cache_words)
#endif
#ifndef CMX_n_contacts_per_pair
#define CMX_n_contacts_per_pair (c_m.n_contacts_per_pair)
#endif
#ifndef CMX_n_roots
#define CMX_n_roots (c_m.n_roots)
#endif
#ifndef CMX_noslip_iterations
#define CMX_noslip_iterations (c_m.noslip_iterations)
#endif
#ifndef CMX_noslip_tolerance
#define CMX_noslip_tolerance (c_m.noslip_tolerance)
#endif
#ifndef CMX_pair_idx
#define CMX_pair_idx (c_m.pair_idx)
#endif
#ifndef CMX_pairs
#define CMX_pairs (c_m.pairs)
#endif
#ifndef CMX_phase_barriers
#define CMX_phase_barriers (c_m.phase_barriers)
#endif
#ifndef CMX_qpos0
#define CMX_qpos0 (c_m.qpos0)
#endif
#ifndef CMX_solver_type
#define CMX_solver_type (c_m.solver_type)
#endif
#ifndef CMX_sort_contacts
#define CMX_sort_contacts (c_m.sort_contacts)
#endif
#ifndef CMX_support_v
#define CMX_support_v (c_m.support_v)
#endif
#ifndef CMX_support_vid
#define CMX_support_vid (c_m.support_vid)
#endif
#ifndef CMX_terrain_cols
#define CMX_terrain_cols (c_m.terrain_cols)
#endif
#ifndef CMX_terrain_hf
#define CMX_terrain_hf (c_m.terrain_hf)
#endif
#ifndef CMX_terrain_hs
#define CMX_terrain_hs (c_m.terrain_hs)
#endif
#ifndef CMX_terrain_maxmin
#define CMX_terrain_maxmin (c_m.terrain_maxmin)
#endif
#ifndef CMX_terrain_rows
#define CMX_terrain_rows (c_m.terrain_rows)
#endif
#ifndef CMX_tolerance
#define CMX_tolerance (c_m.tolerance)
#endif
#ifndef CMX_tri
#define CMX_tri (c_m.tri)
#endif
#ifndef CMX_use_gjk
#define CMX_use_gjk (c_m.use_gjk)
#endif
#ifndef CMX_verts
#define CMX_verts (c_m.verts)
#endif
