// TEST INFRASTRUCTURE ONLY: the two settings TrajectoryManager's constructor reads (FL/TrajectoryManager.cpp:15-16), settable from the wrapper.
#pragma once
struct GlobalAppState {
    unsigned int s_topNActive = 10; float s_minPoseDistSqrt = 0.0f;
    static GlobalAppState& get() { static GlobalAppState s; return s; }
};
