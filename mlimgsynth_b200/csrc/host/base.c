#include "base.h"
#include <time.h>

static _Thread_local char g_err[512];
int g_log_level = LOG_WARN;

void mlis_err_set(const char* fmt, ...)
{
	va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
	mlis_log(LOG_ERROR, "%s", g_err);
}
const char* mlis_err_get(void) { return g_err; }

void mlis_log(int lvl, const char* fmt, ...)
{
	if (lvl > g_log_level) return;
	va_list ap; va_start(ap, fmt);
	fputs(lvl <= LOG_ERROR ? "[MLIS-B200] ERROR " : lvl <= LOG_WARN ? "[MLIS-B200] WARN  " : "[MLIS-B200] ", stderr);
	vfprintf(stderr, fmt, ap);
	fputc('\n', stderr);
	va_end(ap);
}

double time_now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + ts.tv_nsec * 1e-9;
}
