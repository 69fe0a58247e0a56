/* Plain-C consumer of include/isg.h: proves the boundary is a C ABI (no C++ types, no torch types, no name mangling)
 * by compiling against the header with a C99 compiler, linking libisg.so and calling the entry points that need no
 * device: version, error strings and argument validation (every entry point validates before it touches the GPU).
 * Built and run by tests/test_abi.py::test_header_is_c99_and_links_from_c. */
#include <stdio.h>
#include <string.h>

#include "isg.h"

#define CHECK(cond)                                              \
  do {                                                           \
    if (!(cond)) {                                               \
      fprintf(stderr, "abi_check: %s failed (line %d)\n", #cond, __LINE__); \
      return 1;                                                  \
    }                                                            \
  } while (0)

int main(void) {
  CHECK(isg_version() >= 100);
  CHECK(strcmp(isg_error_string(ISG_OK), "ok") == 0);
  CHECK(strstr(isg_error_string(ISG_EINVAL), "ISG_EINVAL") != NULL);
  CHECK(strstr(isg_error_string(ISG_EUNSUPPORTED), "ISG_EUNSUPPORTED") != NULL);
  /* negative edge count: rejected on the host, before any CUDA call */
  CHECK(isg_csr_build(NULL, -1, 4, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL) == ISG_EINVAL);
  /* sizes are pure host arithmetic */
  CHECK(isg_csr_workspace_bytes(16, 64) > 0);
  CHECK(isg_simple_npad(37) == 64);
  printf("abi_check ok: libisg version %d\n", isg_version());
  return 0;
}
