#include "vulkan_core.h"
