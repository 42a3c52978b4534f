/* TEST INFRASTRUCTURE (oracle/ref_host): stand-in for <vulkan/vulkan_core.h>, which is not installed in this image. It declares, as
 * opaque handles / plain integers / empty structs, exactly the Vulkan names that the reference's HOST headers mention
 * (src/core/internal/*.h, src/core/api/*.h, src/core/scene/*.h), so that the reference's own host-side scene-preparation sources
 * (utility/packing.c, scene/transform.c, scene/lighting.c, api/mesh.c ...) compile unmodified from /root/reference with gcc. No Vulkan
 * function is ever called by the code paths the pin tests exercise; the few that are referenced at link time are stubbed to abort in
 * ref_host_entry.c. Nothing here is copied from the Vulkan headers beyond the public type NAMES and the documented layout of
 * VkTransformMatrixKHR (a row-major 3x4 float matrix) and VkExtent2D. */
#ifndef VKRT_ORACLE_VULKAN_SHIM_H
#define VKRT_ORACLE_VULKAN_SHIM_H
#include <stdint.h>
#define VK_DEFINE_SHIM_HANDLE(name) typedef struct name##_T* name;
VK_DEFINE_SHIM_HANDLE(VkInstance) VK_DEFINE_SHIM_HANDLE(VkPhysicalDevice) VK_DEFINE_SHIM_HANDLE(VkDevice) VK_DEFINE_SHIM_HANDLE(VkQueue)
VK_DEFINE_SHIM_HANDLE(VkCommandBuffer) VK_DEFINE_SHIM_HANDLE(VkDeviceMemory) VK_DEFINE_SHIM_HANDLE(VkImage) VK_DEFINE_SHIM_HANDLE(VkBuffer)
VK_DEFINE_SHIM_HANDLE(VkImageView) VK_DEFINE_SHIM_HANDLE(VkShaderModule) VK_DEFINE_SHIM_HANDLE(VkPipeline) VK_DEFINE_SHIM_HANDLE(VkDescriptorPool)
VK_DEFINE_SHIM_HANDLE(VkDebugUtilsMessengerEXT) VK_DEFINE_SHIM_HANDLE(VkSemaphore) VK_DEFINE_SHIM_HANDLE(VkSwapchainKHR) VK_DEFINE_SHIM_HANDLE(VkSurfaceKHR)
VK_DEFINE_SHIM_HANDLE(VkSampler) VK_DEFINE_SHIM_HANDLE(VkQueryPool) VK_DEFINE_SHIM_HANDLE(VkPipelineLayout) VK_DEFINE_SHIM_HANDLE(VkFence)
VK_DEFINE_SHIM_HANDLE(VkDescriptorSetLayout) VK_DEFINE_SHIM_HANDLE(VkDescriptorSet) VK_DEFINE_SHIM_HANDLE(VkCommandPool)
VK_DEFINE_SHIM_HANDLE(VkAccelerationStructureKHR)
typedef uint32_t VkBool32;
typedef uint32_t VkFlags;
typedef uint64_t VkDeviceSize;
typedef uint64_t VkDeviceAddress;
typedef VkFlags VkBufferUsageFlags, VkMemoryPropertyFlags, VkImageUsageFlags, VkDebugUtilsMessageTypeFlagsEXT;
typedef int32_t VkResult, VkFormat, VkPresentModeKHR, VkImageLayout, VkShaderStageFlagBits, VkRayTracingInvocationReorderModeEXT, VkDynamicState,
    VkDebugUtilsMessageSeverityFlagBitsEXT, VkColorSpaceKHR;
#define VK_TRUE 1u
#define VK_FALSE 0u
#define VK_SUCCESS 0
#define VK_SUBOPTIMAL_KHR 1000001003
#define VK_ERROR_DEVICE_LOST (-4)
#define VK_ERROR_OUT_OF_DATE_KHR (-1000001004)
#define VK_NULL_HANDLE 0
/* flag / enum names the host sources pass to the (stubbed) buffer and image helpers; the values are never interpreted */
enum {
    VK_BUFFER_USAGE_TRANSFER_SRC_BIT = 0x1, VK_BUFFER_USAGE_TRANSFER_DST_BIT = 0x2, VK_BUFFER_USAGE_UNIFORM_BUFFER_BIT = 0x10,
    VK_BUFFER_USAGE_STORAGE_BUFFER_BIT = 0x20, VK_BUFFER_USAGE_INDEX_BUFFER_BIT = 0x40, VK_BUFFER_USAGE_VERTEX_BUFFER_BIT = 0x80,
    VK_BUFFER_USAGE_SHADER_DEVICE_ADDRESS_BIT = 0x20000, VK_BUFFER_USAGE_ACCELERATION_STRUCTURE_BUILD_INPUT_READ_ONLY_BIT_KHR = 0x80000,
    VK_BUFFER_USAGE_ACCELERATION_STRUCTURE_STORAGE_BIT_KHR = 0x100000, VK_BUFFER_USAGE_SHADER_BINDING_TABLE_BIT_KHR = 0x400,
    VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT = 0x1, VK_MEMORY_PROPERTY_HOST_VISIBLE_BIT = 0x2, VK_MEMORY_PROPERTY_HOST_COHERENT_BIT = 0x4,
    VK_QUERY_RESULT_64_BIT = 0x1, VK_QUERY_RESULT_WAIT_BIT = 0x2,
    VK_IMAGE_LAYOUT_UNDEFINED = 0, VK_IMAGE_LAYOUT_GENERAL = 1, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL = 6, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL = 7,
    VK_WHOLE_SIZE_SHIM = 0
};
typedef struct VkExtent2D { uint32_t width, height; } VkExtent2D;
typedef struct VkTransformMatrixKHR { float matrix[3][4]; } VkTransformMatrixKHR;
/* for src/core/scene/exposure.c (the 256 one-texel copies of the auto-exposure probe) */
enum { VK_IMAGE_ASPECT_COLOR_BIT = 1 };
typedef struct VkOffset3D { int32_t x, y, z; } VkOffset3D;
typedef struct VkExtent3D { uint32_t width, height, depth; } VkExtent3D;
typedef struct VkImageSubresourceLayers { VkFlags aspectMask; uint32_t mipLevel, baseArrayLayer, layerCount; } VkImageSubresourceLayers;
typedef struct VkBufferImageCopy {
    VkDeviceSize bufferOffset;
    uint32_t bufferRowLength, bufferImageHeight;
    VkImageSubresourceLayers imageSubresource;
    VkOffset3D imageOffset;
    VkExtent3D imageExtent;
} VkBufferImageCopy;
void vkCmdCopyImageToBuffer(VkCommandBuffer commandBuffer, VkImage srcImage, VkImageLayout srcImageLayout, VkBuffer dstBuffer, uint32_t regionCount,
                            const VkBufferImageCopy* regions);
typedef struct VkSurfaceFormatKHR { VkFormat format; VkColorSpaceKHR colorSpace; } VkSurfaceFormatKHR;
typedef struct VkStridedDeviceAddressRegionKHR { VkDeviceAddress deviceAddress; VkDeviceSize stride, size; } VkStridedDeviceAddressRegionKHR;
typedef struct VkSurfaceCapabilitiesKHR { uint32_t opaque[16]; } VkSurfaceCapabilitiesKHR;
typedef struct VkRayTracingShaderGroupCreateInfoKHR { uint32_t opaque[16]; } VkRayTracingShaderGroupCreateInfoKHR;
typedef struct VkPipelineShaderStageCreateInfo { uint32_t opaque[16]; } VkPipelineShaderStageCreateInfo;
typedef struct VkPipelineDynamicStateCreateInfo { uint32_t opaque[16]; } VkPipelineDynamicStateCreateInfo;
typedef struct VkDebugUtilsMessengerCreateInfoEXT { uint32_t opaque[16]; } VkDebugUtilsMessengerCreateInfoEXT;
typedef struct VkAllocationCallbacks { uint32_t opaque[16]; } VkAllocationCallbacks;
typedef void (*PFN_vkVoidFunction)(void);
typedef PFN_vkVoidFunction PFN_vkGetRayTracingShaderGroupStackSizeKHR, PFN_vkGetRayTracingShaderGroupHandlesKHR, PFN_vkGetBufferDeviceAddressKHR,
    PFN_vkGetAccelerationStructureDeviceAddressKHR, PFN_vkGetAccelerationStructureBuildSizesKHR,
    PFN_vkCreateRayTracingPipelinesKHR, PFN_vkCreateAccelerationStructureKHR, PFN_vkCmdTraceRaysKHR, PFN_vkCmdSetRayTracingPipelineStackSizeKHR,
    PFN_vkCmdEndDebugUtilsLabelEXT, PFN_vkCmdBuildAccelerationStructuresKHR, PFN_vkCmdBeginDebugUtilsLabelEXT;
typedef void (*PFN_vkDestroyAccelerationStructureKHR)(VkDevice, VkAccelerationStructureKHR, const VkAllocationCallbacks*);
/* device calls the compiled host sources mention; defined as stubs in ref_host_entry.c */
VkResult vkWaitForFences(VkDevice device, uint32_t fenceCount, const VkFence* fences, VkBool32 waitAll, uint64_t timeout);
VkResult vkMapMemory(VkDevice device, VkDeviceMemory memory, VkDeviceSize offset, VkDeviceSize size, VkFlags flags, void** data);
void vkUnmapMemory(VkDevice device, VkDeviceMemory memory);
void vkDestroyBuffer(VkDevice device, VkBuffer buffer, const VkAllocationCallbacks* allocator);
void vkFreeMemory(VkDevice device, VkDeviceMemory memory, const VkAllocationCallbacks* allocator);
VkResult vkGetQueryPoolResults(VkDevice device, VkQueryPool pool, uint32_t firstQuery, uint32_t queryCount, size_t dataSize, void* data,
                               VkDeviceSize stride, VkFlags flags);
#endif
