/*
 * Minimal stand-in for <jni.h>, for the build image only (no JDK exists there): exactly the declarations
 * integration/jni/filmyou_rm2_jni.c uses, source-compatible with the C binding of the real header
 * ((*env)->Fn(env, ...) through a function table).  TEST INFRASTRUCTURE; a real build uses $JAVA_HOME/include/jni.h.
 */
#ifndef FY_MOCK_JNI_H
#define FY_MOCK_JNI_H
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL

typedef int32_t jint;
typedef int64_t jlong;
typedef double jdouble;
typedef void* jobject;
typedef jobject jclass;
typedef jobject jstring;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
    void* (*GetDirectBufferAddress)(JNIEnv* env, jobject buf);
    jlong (*GetDirectBufferCapacity)(JNIEnv* env, jobject buf);
    jstring (*NewStringUTF)(JNIEnv* env, const char* utf);
    jclass (*FindClass)(JNIEnv* env, const char* name);
    jint (*ThrowNew)(JNIEnv* env, jclass cls, const char* message);
};
#endif
