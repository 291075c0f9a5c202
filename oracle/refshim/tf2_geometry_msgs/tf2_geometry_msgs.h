#include <tf2/utils.h>
