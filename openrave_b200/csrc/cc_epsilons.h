// cc_epsilons.h — the reference's tolerance constants for dReal = double, in ONE place for the device code, the planner host
// and (by value) openrave_b200/kinbody.py.  g_fEpsilon = 1e-15 (include/openrave/openrave.h:130); the derived ones are
// RavePow(g_fEpsilon, x) (src/libopenrave/libopenrave.h:140-142).  Literals = pow() evaluated in IEEE double.
#pragma once
#define CC_EPSILON 1e-15
#define CC_EPSILON_LINEAR 3.162277660168377e-14             /* pow(1e-15, 0.9)  g_fEpsilonLinear */
#define CC_EPSILON_JOINT_LIMIT 9.999999999999986e-13        /* pow(1e-15, 0.8)  g_fEpsilonJointLimit: a mimic value this close outside a limit snaps onto it */
#define CC_EPSILON_EVAL_JOINT_LIMIT 1.7782794100389216e-10  /* pow(1e-15, 0.65) g_fEpsilonEvalJointLimit: a mimic value farther outside than this is clamped */
