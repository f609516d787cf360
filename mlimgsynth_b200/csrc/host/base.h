/* base.h -- small shared helpers of the B200 host layer (C11). */
#pragma once
#include <stdarg.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* Error convention (same as the reference C API, mlimgsynth.h:68-77): functions return >= 0 on
 * success, < 0 on error; the message is kept per thread and surfaced by mlis_errstr_get(). */
void  mlis_err_set(const char* fmt, ...);
const char* mlis_err_get(void);
#define FAIL(code, ...)  do { mlis_err_set(__VA_ARGS__); return (code); } while (0)
#define CHECK(expr)      do { int r_ = (expr); if (r_ < 0) return r_; } while (0)

enum { LOG_NONE = 0, LOG_ERROR = 10, LOG_WARN = 20, LOG_INFO = 30, LOG_VERBOSE = 40, LOG_DEBUG = 50 };
extern int g_log_level;
void mlis_log(int lvl, const char* fmt, ...);
#define log_info(...)   mlis_log(LOG_INFO, __VA_ARGS__)
#define log_warn(...)   mlis_log(LOG_WARN, __VA_ARGS__)
#define log_debug(...)  mlis_log(LOG_DEBUG, __VA_ARGS__)

double time_now(void);   /* monotonic seconds */

static inline void* xmalloc(size_t n) { void* p = malloc(n ? n : 1); if (!p) { fprintf(stderr, "out of memory\n"); abort(); } return p; }
static inline void* xcalloc(size_t n, size_t s) { void* p = calloc(n ? n : 1, s ? s : 1); if (!p) { fprintf(stderr, "out of memory\n"); abort(); } return p; }
static inline void* xrealloc(void* q, size_t n) { void* p = realloc(q, n ? n : 1); if (!p) { fprintf(stderr, "out of memory\n"); abort(); } return p; }
static inline char* xstrdup(const char* s) { size_t n = strlen(s) + 1; char* p = xmalloc(n); memcpy(p, s, n); return p; }

/* growable array of T: arr_push(&ptr, &count, &cap, value) */
#define ARR_PUSH(ptr, n, cap, val) do { \
	if ((n) == (cap)) { (cap) = (cap) ? (cap) * 2 : 16; (ptr) = xrealloc((ptr), (size_t)(cap) * sizeof(*(ptr))); } \
	(ptr)[(n)++] = (val); } while (0)
