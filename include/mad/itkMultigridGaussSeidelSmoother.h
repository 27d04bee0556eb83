/*
 * mad/itkMultigridGaussSeidelSmoother.h -- forwarding header, so that user code written for the reference
 * (#include "mad/itkMultigridGaussSeidelSmoother.h", /root/reference/test/*.cxx) compiles unchanged against the B200 drop-in:
 * the class name survives as a tag type (mad/itkMultigridSmootherTags.h), the smoother itself runs inside libmadgpu.so.
 */
#ifndef __itkMultigridGaussSeidelSmoother_h
#define __itkMultigridGaussSeidelSmoother_h
#include "itkMultigridSmootherTags.h"
#endif
