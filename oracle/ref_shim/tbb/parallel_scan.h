// TEST INFRASTRUCTURE ONLY (oracle build shim)
#pragma once
#include "blocked_range.h"
