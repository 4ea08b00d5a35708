// Which GPU the objects created by the calling thread use (the reference picks devices.at(0),
// viso/opencl_wrapper.cpp:89; sharded runs put one sequence set on each GPU).
#ifndef VISOB_DEVICE_H
#define VISOB_DEVICE_H
namespace visob {
void set_device(int device);     // affects Matcher / filter:: objects created afterwards by this thread
int current_device();
void set_pipeline_depth(int depth);   // steps the sequence runner keeps in flight per sequence (1 .. 3, default 3; VISOB_DEPTH)
int pipeline_depth();
}
#endif
