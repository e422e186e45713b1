// Test-only: the C++ shim must compile against include/lldba.h and link against liblldba.so.
#include <algorithm>
#include "../../lld_slam_b200/host/lld_shim.h"
int main() {
  lld::LocalWindow w;
  lld::LocalBAResult r;
  (void)w; (void)r;
  uint8_t a[32] = {0}, b[32] = {0};
  b[3] = 0x0F;
  return lld::ORBmatcher::DescriptorDistance(a, b) == 4 ? 0 : 1;
}
