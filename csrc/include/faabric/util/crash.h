#pragma once

#include <initializer_list>

namespace faabric::util {

// Installs a backtrace-printing handler for fatal signals.  SIGSEGV is left
// alone by default because the segfault dirty tracker owns it.
void setUpCrashHandler(int sig = -1);

void printStackTrace(void* contextR = nullptr);

}
