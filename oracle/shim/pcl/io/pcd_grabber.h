// TEST SCAFFOLDING: the reference includes this header but uses nothing from it.
#pragma once
