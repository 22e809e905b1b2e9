#pragma once
#include "mlib_min.h"
