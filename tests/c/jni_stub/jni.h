/* minimal stand-in for <jni.h>: only the types and function-table entries integration/jni/sfgpu_jni.c uses, so that the adapter's calls into include/sfgpu.h are type-checked on a box without a JDK (tests/test_abi.py). Not a JNI implementation. */
#include <stdint.h>
typedef int32_t jint; typedef int64_t jlong; typedef int8_t jbyte; typedef double jdouble; typedef uint8_t jboolean; typedef jint jsize;
typedef void *jobject; typedef jobject jclass, jstring, jarray, jobjectArray, jbyteArray, jintArray, jlongArray, jdoubleArray;
#define JNIEXPORT
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_; typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
  void *(*GetDirectBufferAddress)(JNIEnv*, jobject); jobject (*NewDirectByteBuffer)(JNIEnv*, void*, jlong); jstring (*NewStringUTF)(JNIEnv*, const char*);
  void (*GetDoubleArrayRegion)(JNIEnv*, jdoubleArray, jsize, jsize, jdouble*); void (*SetDoubleArrayRegion)(JNIEnv*, jdoubleArray, jsize, jsize, const jdouble*);
  void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, const jlong*); jobject (*GetObjectArrayElement)(JNIEnv*, jobjectArray, jsize);
  jbyte *(*GetByteArrayElements)(JNIEnv*, jbyteArray, jboolean*); jint *(*GetIntArrayElements)(JNIEnv*, jintArray, jboolean*); jdouble *(*GetDoubleArrayElements)(JNIEnv*, jdoubleArray, jboolean*);
  void (*ReleaseByteArrayElements)(JNIEnv*, jbyteArray, jbyte*, jint); void (*ReleaseIntArrayElements)(JNIEnv*, jintArray, jint*, jint); void (*ReleaseDoubleArrayElements)(JNIEnv*, jdoubleArray, jdouble*, jint);
  jsize (*GetArrayLength)(JNIEnv*, jarray);
};
