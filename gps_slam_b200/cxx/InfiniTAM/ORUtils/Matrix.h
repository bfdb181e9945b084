// forwards to the gps_slam_b200 facade of the InfiniTAM interface (the reference has a header of this name and path under
// InfiniTAM/; its callers keep their #include lines)
#pragma once
#include "../gsb_itm.h"
