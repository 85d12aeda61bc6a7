// stand-in of ocs2_sqp/SqpSettings.h: the fields the adapter reads (task.info:79-96)
#pragma once
#include "ocs2_core/Types.h"
namespace ocs2 { namespace sqp { struct Settings { scalar_t dt = 0.01; size_t sqpIteration = 1; }; } }
