/* permonqppf.h -- compatibility name: reference code that includes <permonqppf.h> gets the B200 C ABI. */
#pragma once
#include "permon_b200.h"
