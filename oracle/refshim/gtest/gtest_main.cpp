// main() for the reference's test files compiled against the gtest stand-in (TEST INFRASTRUCTURE ONLY).
#include <gtest/gtest.h>

int
main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
