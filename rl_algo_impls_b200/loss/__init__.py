from .teacher_kl_loss import TeacherKLLoss

__all__ = ["TeacherKLLoss"]
