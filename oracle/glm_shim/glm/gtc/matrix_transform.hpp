// glm-compatible shim (oracle/ test infrastructure) — the reference includes this header
// (camera.hpp:4) but uses nothing from it.
#pragma once
#include "../glm.hpp"
