/* permonqp.h -- compatibility name: reference code that includes <permonqp.h> gets the B200 C ABI. */
#pragma once
#include "permon_b200.h"
