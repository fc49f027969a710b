// TEST SCAFFOLDING (oracle build only) -- not product code.
// FragmentOptimizer's TUs include "StdAfx.h"; the file in the reference tree is "stdafx.h".  Adds the libc
// headers MSVC pulled in implicitly and MSVC's _isnan (FragmentOptimizer/PointCloud.cpp:30).
#pragma once
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <string>
#include <vector>
#include <iostream>
#ifndef _isnan
#define _isnan(x) std::isnan(x)
#endif
#include "stdafx.h"
