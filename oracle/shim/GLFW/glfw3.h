// Stand-in for GLFW: only the opaque window type is named by camera.h:20.
#pragma once
struct GLFWwindow;
