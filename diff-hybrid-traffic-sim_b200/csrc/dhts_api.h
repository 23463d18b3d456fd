// Internal: pulls in the public C ABI and defines the export attribute.
#pragma once
#include "../../include/dhts.h"
#define DHTS_EXPORT extern "C" __attribute__((visibility("default")))
