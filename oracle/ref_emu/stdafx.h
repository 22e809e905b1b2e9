// TEST INFRASTRUCTURE ONLY: empty stand-in for the reference's Windows / precompiled header <stdafx.h>.
#pragma once
