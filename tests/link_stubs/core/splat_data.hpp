// stub: see camera.hpp
#pragma once
