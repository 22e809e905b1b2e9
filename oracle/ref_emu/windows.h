// TEST INFRASTRUCTURE ONLY: stand-in for <windows.h> (the reference's headers rely on what it drags in).
#pragma once
#include <algorithm>
#include <cfloat>
#include <fstream>
#include <list>
#include <mutex>
#include <string>
