"""Oracle (TEST INFRASTRUCTURE ONLY — never imported by the product): a literal Python restatement of the loop
body of the reference's Composer::pose_assember_thread (reference src/Composer.cpp:24-209), one keyframe after
the other with 4x4 numpy matrices and the generic inverse, exactly as the reference walks it — including the
`jmb[world].rbegin()` look-ups that make dead-zone keyframes depend on what was assembled before them.

Pinned to the reference's own Composer::pose_assember_thread (compiled unmodified into oracle/_ref/libref_frontend.so and
single-stepped) by tests/test_reference_frontend.py::test_pose_assembly_matches_the_reference_composer_thread, over the
states a session goes through; known-answer cases in tests/test_composer.py."""
import numpy as np


def assemble(manager, slam_poses, solved_until):
    """manager: oracle.frontend.Manager (or anything with .poses [4x4], .stamps, which_world_is_this,
    nodeidx_of_world_i_ended, .worlds); slam_poses: list of 4x4 optimised poses (slam->getNodePose(i) for
    i < len); solved_until: slam->solvedUntil().
    Returns (lmb [n][4][4], world_id [n], jmb {world: [4x4,...]})."""
    n = len(manager.poses)
    jmb, lmb, wid = {}, [], []
    if n == 0:                                                       # Composer.cpp:26-30
        return np.zeros((0, 4, 4)), np.zeros(0, int), jmb
    su = solved_until                                                # :35
    su_world = manager.which_world_is_this(manager.stamps[su])       # :36
    n_slam = len(slam_poses)
    W = manager.worlds
    for i in range(n):                                               # :55
        world_id = manager.which_world_is_this(manager.stamps[i])    # :57
        set_id = W.find_setID_of_world_i(world_id)                   # :58
        if i <= su:                                                  # :63
            if world_id >= 0:                                        # :69
                T = slam_poses[i] if i < n_slam else manager.poses[i]                     # :73-84
            else:                                                    # :89-101
                last_idx = manager.nodeidx_of_world_i_ended(-world_id - 1)
                w_T_last = jmb[-world_id - 1][-1]
                T = w_T_last @ (np.linalg.inv(manager.poses[last_idx]) @ manager.poses[i])
        else:                                                        # :118
            last_idx, from_mgr = -1, -1
            if su == 0:                                              # :128-132
                T = manager.poses[i]; from_mgr = 0
            elif world_id >= 0 and su_world == world_id:             # :134-135
                last_idx = su
            elif world_id >= 0 and su_world != world_id:             # :137-139
                T = manager.poses[i]
            else:                                                    # :140-146 (world_id < 0)
                li = manager.nodeidx_of_world_i_ended(-world_id - 1)
                T = jmb[-world_id - 1][-1] @ (np.linalg.inv(manager.poses[li]) @ manager.poses[i])
            if last_idx >= 0:                                        # :155-165
                w_T_last = slam_poses[last_idx] if last_idx < n_slam else manager.poses[last_idx]
                T = w_T_last @ (np.linalg.inv(manager.poses[last_idx]) @ manager.poses[i])
            if world_id != set_id and from_mgr == 0:                 # :172-190
                if W.is_exist(set_id, world_id):
                    T = W.getPoseBetweenWorlds(set_id, world_id) @ T
        jmb.setdefault(world_id, []).append(T)                       # :192-200
        lmb.append(T); wid.append(world_id)
    return np.array(lmb), np.array(wid), jmb
