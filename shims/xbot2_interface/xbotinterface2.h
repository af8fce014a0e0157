// Oracle build shim (test infrastructure): the reference includes this header for Eigen
// and the logger types; the robot model itself is never touched on the solver path.
#pragma once
#include <Eigen/Dense>
#include <memory>
#include <xbot2_interface/logger.h>
#include <matlogger2/matlogger2.h>
namespace XBot { inline namespace v2 { class ModelInterface; } }
