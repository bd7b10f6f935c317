// Library-wide state of libsagan_b200.so: error text, ABI version, launch counter.
#include <stdarg.h>

#include "common.cuh"

namespace sagan {

std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_deterministic_forward{0};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

void set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
}

}  // namespace sagan

extern "C" int sagan_abi_version(void) { return SAGAN_B200_ABI_VERSION; }
extern "C" const char* sagan_last_error(void) { return sagan::err_buf(); }
extern "C" unsigned long long sagan_launch_count(void) { return sagan::g_launches.load(); }
extern "C" int sagan_deterministic_forward(int set) {
  if (set == 0 || set == 1) sagan::g_deterministic_forward.store(set);
  return sagan::g_deterministic_forward.load();
}
