/* Host shim for the one Pico-SDK header the reference's compute path includes
 * (src/components/correlations.h:3).  TEST INFRASTRUCTURE ONLY.
 * The clock is injected by the driver (oracle/ref_driver.c). */
#pragma once
#include <stdint.h>
typedef uint64_t absolute_time_t;
absolute_time_t get_absolute_time(void);
