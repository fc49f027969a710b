// TEST SCAFFOLDING (oracle build only) -- not product code.
// The reference TUs include "StdAfx.h" (capital S/A) while the file shipped in
// /root/reference/Integrate is "stdafx.h"; on a case-sensitive file system we
// forward to it and add the libc/libstdc++ headers MSVC pulled in implicitly.
#pragma once
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <string>
#include <vector>
#include <iostream>
#include "stdafx.h"
