/* TEST INFRASTRUCTURE: stand-in for <GLFW/glfw3.h> (see ../vulkan/vulkan_core.h). The reference's internal state struct holds a window pointer. */
#ifndef VKRT_ORACLE_GLFW_SHIM_H
#define VKRT_ORACLE_GLFW_SHIM_H
typedef struct GLFWwindow GLFWwindow;
#endif
