// costmap_2d/cost_values.h — stand-in (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md)
#ifndef DPOSE_REFSHIM_COST_VALUES
#define DPOSE_REFSHIM_COST_VALUES
namespace costmap_2d {
static const unsigned char NO_INFORMATION = 255;
static const unsigned char LETHAL_OBSTACLE = 254;
static const unsigned char INSCRIBED_INFLATED_OBSTACLE = 253;
static const unsigned char FREE_SPACE = 0;
}  // namespace costmap_2d
#endif
