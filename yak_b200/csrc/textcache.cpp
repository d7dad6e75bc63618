// textcache.cpp - see textcache.h
#include "textcache.h"
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <mutex>

namespace yakb {

namespace {
std::mutex g_mu;
std::string g_key;     // identity of the cached file: real path | size | mtime
int g_fd = -1;
uint64_t g_bytes = 0;

std::string identity(const char *fn)
{
	struct stat st;
	if (fn == nullptr || strcmp(fn, "-") == 0 || stat(fn, &st) != 0 || !S_ISREG(st.st_mode)) return "";
	char *rp = realpath(fn, nullptr);
	std::string k = rp ? rp : fn;
	free(rp);
	k += "|" + std::to_string((long long)st.st_size) + "|" + std::to_string((long long)st.st_mtim.tv_sec) + "." + std::to_string((long)st.st_mtim.tv_nsec);
	return k;
}
}

uint64_t text_cache_budget()
{
	const char *e = getenv("YAKB_TEXT_CACHE_GB");
	const double gb = e ? atof(e) : 0.0;
	return gb > 0 ? (uint64_t)(gb * (double)(1ull << 30)) : 0;
}

int text_cache_begin(const char *fn, std::string *key)
{
	if (text_cache_budget() == 0) return -1;
	*key = identity(fn);
	if (key->empty()) return -1;
	return memfd_create("yakb_text", MFD_CLOEXEC);
}

void text_cache_end(const std::string &key, int fd, uint64_t bytes)
{
	if (fd < 0) return;
	if (bytes == UINT64_MAX || key.empty()) { close(fd); return; }
	std::lock_guard<std::mutex> lk(g_mu);
	if (g_fd >= 0) close(g_fd); // one file at a time
	g_fd = fd; g_key = key; g_bytes = bytes;
}

std::string text_cache_lookup(const char *fn)
{
	const std::string k = identity(fn);
	std::lock_guard<std::mutex> lk(g_mu);
	if (g_fd < 0 || k.empty() || k != g_key) return "";
	return "/proc/self/fd/" + std::to_string(g_fd);
}

void text_cache_release()
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (g_fd >= 0) close(g_fd);
	g_fd = -1; g_key.clear(); g_bytes = 0;
}

} // namespace yakb
