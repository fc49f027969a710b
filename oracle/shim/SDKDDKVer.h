// TEST SCAFFOLDING: empty stand-in for the Windows SDK header named by targetver.h.
#pragma once
