// placeholder for the reference header of this name: image / IMU sources and the video writer are not on the per-frame path
// (SURVEY.md section 8 marks them out of scope); only the namespace exists so that `using namespace InputSource;` compiles
#pragma once
#include "../gsb_itm.h"
namespace InputSource
{
class IMUSourceEngine;
}
