// stub: the real core/camera.hpp belongs to the reference product and is not needed by rasterizer_autograd.cpp
#pragma once
