// misc.cu -- library version.
#include "../../include/coocc_b200.h"
extern "C" int coocc_version(void) { return 100; }
