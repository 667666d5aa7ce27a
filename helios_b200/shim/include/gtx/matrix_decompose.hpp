// gtx/matrix_decompose.hpp — glm::decompose lives in gtc/quaternion.hpp of the GLM stand-in (see ../glm.hpp)
#pragma once
#include <gtc/quaternion.hpp>
