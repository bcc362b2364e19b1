// userfaultfd definitions.  The reference vendors the kernel's header
// (include/faabric/util/userfaultfd.h) because its build image lacked some of
// them; this image's <linux/userfaultfd.h> has everything the uffd dirty
// trackers use (src/util/dirty.cpp).
#pragma once

extern "C"
{
#include <linux/userfaultfd.h>
}
