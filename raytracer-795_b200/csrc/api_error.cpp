// Thread-local last-error string behind rt_last_error() (include/rt795.h).
#include "../../include/rt795.h"
#include <string>

namespace rt795 {
static thread_local std::string g_last_error;
void SetLastError(const std::string& msg) { g_last_error = msg; }
}  // namespace rt795

extern "C" {
const char* rt_last_error(void) { return rt795::g_last_error.c_str(); }
int rt_abi_version(void) { return RT795_ABI_VERSION; }
}
