/*
 * mad/itkMultigridWeightedJacobiSmoother.h -- forwarding header, so that user code written for the reference
 * (#include "mad/itkMultigridWeightedJacobiSmoother.h", /root/reference/test/*.cxx) compiles unchanged against the B200 drop-in:
 * the class name survives as a tag type (mad/itkMultigridSmootherTags.h), the smoother itself runs inside libmadgpu.so.
 */
#ifndef __itkMultigridWeightedJacobiSmoother_h
#define __itkMultigridWeightedJacobiSmoother_h
#include "itkMultigridSmootherTags.h"
#endif
