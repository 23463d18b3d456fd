#include "dhts_api.h"
DHTS_EXPORT int dhts_version(void) { return 100; }
