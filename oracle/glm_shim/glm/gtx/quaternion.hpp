// part of the minimal GLM-compatible shim (see ../glm.hpp)
#pragma once
#include "../glm.hpp"
