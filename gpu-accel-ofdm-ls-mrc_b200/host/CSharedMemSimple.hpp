// CSharedMemSimple.hpp -- POSIX shared-memory segment, same class and methods as the
// reference's CSharedMemSimple (CSharedMemSimple.hpp:70-140): open-or-create the named
// object read/write, size it, map it shared; the master unmaps and unlinks on destruction.
// Written from scratch for the B200 host side.  Differences, all deliberate:
//   * sizes are size_t (the reference takes unsigned int -> segments < 4 GiB, :88);
//   * needs no HAVE_UNISTD_H define (the reference includes <io.h> without it, :34-38);
//   * the unused page/notebook structs (:42-68) are gone.
// Failure behaviour is the reference's: perror + exit (:78-84, :103-104).
#ifndef LSMRC_HOST_CSHAREDMEMSIMPLE_HPP_
#define LSMRC_HOST_CSHAREDMEMSIMPLE_HPP_

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <string>

class CSharedMemSimple {
   public:
    CSharedMemSimple(std::string shm_uid, size_t sizeInBytes) : name_(shm_uid), bytes_(sizeInBytes)
    {
        fd_ = shm_open(name_.c_str(), O_CREAT | O_RDWR, S_IRUSR | S_IWUSR);
        if (fd_ == -1) die("open");
        if (ftruncate(fd_, (off_t)bytes_) == -1) die("ftruncate");
        base_ = mmap(nullptr, bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd_, 0);
        if (base_ == MAP_FAILED) exit(-1);
    }
    CSharedMemSimple(const CSharedMemSimple&) = delete;
    CSharedMemSimple& operator=(const CSharedMemSimple&) = delete;

    ~CSharedMemSimple()
    {
        if (base_ && base_ != MAP_FAILED) munmap(base_, bytes_);
        if (fd_ != -1) close(fd_);
        if (master_) shm_unlink(name_.c_str());
    }

    void set_master_mode() { master_ = true; }
    size_t nBytes() { return bytes_; }
    void* ptr() { return base_; }
    void info()
    {
        printf("SHM info: %s, %s\n", name_.c_str(), master_ ? "Master" : "Slave");
        printf("SHM bytes allocated: %zu\n", bytes_);
    }

   private:
    [[noreturn]] void die(const char* what)
    {
        perror(what);
        exit(EXIT_FAILURE);
    }
    std::string name_;
    size_t bytes_ = 0;
    void* base_ = nullptr;
    int fd_ = -1;
    bool master_ = false;
};

#endif
