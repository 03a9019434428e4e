// Stand-in for <GL/gl.h>: cuda_gl_interop.h and the reference's interop.h only need the basic GL typedefs.
#pragma once
typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef int GLint;
typedef int GLsizei;
typedef unsigned int GLbitfield;
typedef float GLfloat;
typedef unsigned char GLboolean;
typedef void GLvoid;
