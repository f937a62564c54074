/* empty host stub for the Pico SDK header of the same name (test infrastructure) */
#pragma once
#include "../sc_sdk.h"
