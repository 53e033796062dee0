// see ../extension.h (TEST INFRASTRUCTURE ONLY)
#pragma once
#include "../extension.h"
