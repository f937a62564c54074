/* sc_sdk.h -- TEST INFRASTRUCTURE.  The handful of Pico-SDK identifiers the reference's sample_compute.h and its
 * vendored protothread header touch, so that the reference's OWN orchestration code (protothread_sample_and_compute,
 * src/sample_compute.h:45-150) runs on the host against either the reference objects or libat_b200.so.
 * busy_wait_until() is the hook that advances a recorded ADC stream (see oracle/sc_host.c). */
#pragma once
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>

typedef uint64_t absolute_time_t;
typedef volatile uint32_t spin_lock_t;
typedef struct uart_inst uart_inst_t;
#define uart0 ((uart_inst_t *)0)

absolute_time_t get_absolute_time(void);
uint64_t time_us_64(void);
void busy_wait_until(absolute_time_t t);
static inline absolute_time_t delayed_by_us(absolute_time_t t, uint64_t us) { return t + us; }
static inline void gpio_put(unsigned pin, bool v) { (void)pin; (void)v; }
static inline unsigned get_core_num(void) { return 0; }
static inline void uart_putc(uart_inst_t *u, char c) { (void)u; (void)c; }
static inline char uart_getc(uart_inst_t *u) { (void)u; return 0; }
static inline bool uart_is_writable(uart_inst_t *u) { (void)u; return true; }
static inline bool uart_is_readable(uart_inst_t *u) { (void)u; return false; }
