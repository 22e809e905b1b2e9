// TEST INFRASTRUCTURE ONLY: empty stand-in (TrajectoryManager.h includes it, uses nothing from it)
#pragma once
